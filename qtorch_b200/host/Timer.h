// qtorch_b200/host/Timer.h -- wall + CPU stopwatch with the interface of /root/reference/src/Timer.h:25-60
// (public fields included; getElapsed() reports 0 until start() is called, which the planners rely on:
// an un-armed watchdog never fires, Network.h:899).
#pragma once
#include <chrono>
#include <ctime>

namespace qtorch {

class Timer {
public:
    void start() {
        mStart = std::chrono::high_resolution_clock::now();
        mCPUClockStart = std::clock();
        mStarted = true;
    }
    void reset() { mStarted = false; }
    double getElapsed() {
        if (!mStarted) return 0.0;
        const auto dt = std::chrono::high_resolution_clock::now() - mStart;
        return std::chrono::duration_cast<std::chrono::nanoseconds>(dt).count() * 1e-9;
    }
    double getCPUElapsed() { return static_cast<double>(std::clock() - mCPUClockStart) / CLOCKS_PER_SEC; }

    std::chrono::high_resolution_clock::time_point mStart;
    std::clock_t mCPUClockStart{0};
    bool mStarted{false};
};

}  // namespace qtorch
