// wall clock in seconds (the reference times mgtest with gettimeofday, dtime.hpp:5-11)
#pragma once
#include <chrono>
inline double dtime() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
