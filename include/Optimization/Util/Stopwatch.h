// B200 build of the drop-in header layer: same include path and names as the
// reference (include/Optimization/Util/Stopwatch.h:15-29), written from scratch.
#pragma once
#include <chrono>

namespace Optimization {
namespace Stopwatch {

using clock_type = std::chrono::high_resolution_clock;

inline clock_type::time_point tick() { return clock_type::now(); }

// Seconds since `start`, truncated to whole milliseconds exactly like the reference
// (its duration_cast<milliseconds>, Stopwatch.h:26-28): TNT's ElapsedTime test has 1 ms grain.
inline double tock(const clock_type::time_point &start) {
  const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(clock_type::now() - start);
  return ms.count() / 1000.0;
}

}  // namespace Stopwatch
}  // namespace Optimization
