// stand-in for cho::util::UTimer (stopwatch used for printing only, align_icp.cpp:81,87,93)
#pragma once
namespace cho { namespace util {
class UTimer { public: explicit UTimer(bool = false) {} long StopAndGetElapsedTime() { return 0; } };
}}
