// Drop-in header layer (reference: include/Optimization/Base/Concepts.h:37-88).
// Field names, defaults and meaning are the reference's; nothing else is shared.
#pragma once
#include <functional>
#include <limits>
#include <vector>

namespace Optimization {

template <typename Variable, typename Scalar = double, typename... Args>
using Objective = std::function<Scalar(const Variable &X, Args &...args)>;

struct OptimizerParams {
  size_t max_iterations = 100;
  double max_computation_time = std::numeric_limits<double>::max();
  bool log_iterates = false;
  bool verbose = false;
  size_t precision = 3;   // digits printed when verbose
};

template <typename Variable, typename Scalar = double>
struct OptimizerResult {
  Variable x;
  Scalar f;
  double elapsed_time;
  std::vector<Scalar> objective_values;
  std::vector<double> time;
  std::vector<Variable> iterates;
};

}  // namespace Optimization
