#!/bin/bash
# time every experiment variant of the library (ar_voxel_project_b200/lib/variants/*.so, built with build.build(out=...))
for c in ${CONFIGS:-C4 C5}; do
echo "== base $c"; python tools/profile_carve.py --config $c --reps 6 2>&1 | grep "carve ms\|executed" | tail -4
for f in ar_voxel_project_b200/lib/variants/*.so; do echo "== $f $c"; python tools/profile_carve.py --config $c --reps 6 --lib $f 2>&1 | grep "carve ms\|executed" | tail -4; done
done
