// Bucket planner: groups parameter tensors into flat gradient buckets.
// Contract mirrors what DDP's construction path computes (ref: ddp_example.py:64 →
// torch/nn/parallel/distributed.py:1224-1280, c10d reducer.hpp:592-598): tensors are visited in
// the given order, grouped by (dtype, device), a bucket closes when it reaches the current size
// limit of its group, limits advance through `size_limits` and the last one sticks.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace pdt {

struct PlanInput {
  int64_t nbytes;
  int64_t group_key;  // encodes (dtype, device)
};

struct PlanResult {
  std::vector<std::vector<int64_t>> buckets;   // indices into the input order
  std::vector<int64_t> size_limits;            // limit in force when each bucket closed
};

// order: visit order (empty → 0..n-1). When `order` is given (rebuild from observed grad-ready
// order) buckets keep their creation order; otherwise they are sorted by smallest member index.
PlanResult plan_buckets(const std::vector<PlanInput>& tensors, const std::vector<int64_t>& size_limits,
                        const std::vector<int64_t>& order);

}  // namespace pdt
