#include "bucket_plan.h"

#include <algorithm>
#include <map>
#include <stdexcept>

namespace pdt {

PlanResult plan_buckets(const std::vector<PlanInput>& tensors, const std::vector<int64_t>& size_limits,
                        const std::vector<int64_t>& order) {
  if (size_limits.empty()) throw std::invalid_argument("plan_buckets: need at least one size limit");
  if (!order.empty() && order.size() != tensors.size())
    throw std::invalid_argument("plan_buckets: order must cover every tensor exactly once");
  struct Open {
    std::vector<int64_t> idx;
    int64_t bytes = 0;
  };
  std::map<int64_t, Open> open;           // per group
  std::map<int64_t, size_t> limit_pos;    // per group: index into size_limits
  PlanResult res;
  const size_t n = tensors.size();
  for (size_t v = 0; v < n; ++v) {
    int64_t i = order.empty() ? static_cast<int64_t>(v) : order[v];
    if (i < 0 || static_cast<size_t>(i) >= n) throw std::out_of_range("plan_buckets: bad index in order");
    const PlanInput& t = tensors[static_cast<size_t>(i)];
    Open& o = open[t.group_key];
    o.idx.push_back(i);
    o.bytes += t.nbytes;
    size_t& lp = limit_pos[t.group_key];  // value-initialised to 0 on first touch
    int64_t limit = size_limits[lp];
    if (o.bytes >= limit) {
      res.buckets.push_back(std::move(o.idx));
      res.size_limits.push_back(limit);
      open.erase(t.group_key);
      if (lp + 1 < size_limits.size()) ++lp;
    }
  }
  for (auto& kv : open) {
    if (kv.second.idx.empty()) continue;
    res.buckets.push_back(std::move(kv.second.idx));
    res.size_limits.push_back(size_limits[limit_pos[kv.first]]);
  }
  if (order.empty()) {
    // stable order by smallest member so the layout is independent of map iteration order
    std::vector<size_t> perm(res.buckets.size());
    for (size_t k = 0; k < perm.size(); ++k) perm[k] = k;
    std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
      return *std::min_element(res.buckets[a].begin(), res.buckets[a].end()) <
             *std::min_element(res.buckets[b].begin(), res.buckets[b].end());
    });
    PlanResult sorted;
    for (size_t k : perm) {
      sorted.buckets.push_back(std::move(res.buckets[k]));
      sorted.size_limits.push_back(res.size_limits[k]);
    }
    return sorted;
  }
  return res;
}

}  // namespace pdt
