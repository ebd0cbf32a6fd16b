#!/bin/bash
# time every experiment variant of the library (ar_voxel_project_b200/lib/variants/*.so, built with build.build(out=...)) on C4
echo "== base"; python tools/profile_carve.py --config C4 --reps 6 | grep "carve ms" | tail -4
for f in ar_voxel_project_b200/lib/variants/*.so; do echo "== $f"; python tools/profile_carve.py --config C4 --reps 6 --lib $f | grep "carve ms" | tail -4; done
