// stand-in for the reference's Benchmark.h: the singleton's phase hooks (Benchmark.h:59-63,86-124), declarations only
#pragma once
class Benchmark {
   public:
    static Benchmark& GetInstance();
    void LogCarving(bool start);
    void LogColoring(bool start);
    void LogPostProcessing(bool start);
    void LogMarchingCubes(bool start);
    void LogOverall(bool start);
};
