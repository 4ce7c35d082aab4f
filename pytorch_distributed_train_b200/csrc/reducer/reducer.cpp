#include "reducer.h"

#include <c10/util/Exception.h>
#include <torch/csrc/autograd/engine.h>
#include <torch/csrc/autograd/functions/accumulate_grad.h>
#include <torch/csrc/autograd/utils/lambda_post_hook.h>
#include <torch/csrc/autograd/variable.h>
#include <torch/csrc/autograd/python_variable.h>
#include <torch/csrc/utils/pybind.h>

#include <deque>
#include <unordered_set>

namespace pdt {

namespace {
constexpr int64_t kAlignBytes = 16;  // keep every slot float4-addressable for the fused kernels

int64_t group_key(const at::Tensor& t) {
  return (static_cast<int64_t>(t.scalar_type()) << 20) | (static_cast<int64_t>(t.device().type()) << 12) |
         static_cast<int64_t>(t.device().has_index() ? t.device().index() + 1 : 0);
}

bool same_view(const at::Tensor& a, const at::Tensor& b) {
  return a.defined() && b.defined() && a.data_ptr() == b.data_ptr() && a.sizes() == b.sizes() &&
         a.strides() == b.strides() && a.scalar_type() == b.scalar_type();
}

}  // namespace

Reducer::Reducer(std::vector<at::Tensor> params, std::vector<std::vector<int64_t>> bucket_indices,
                 std::shared_ptr<Comm> comm, int64_t bucket_bytes_cap, int64_t first_bucket_bytes_cap,
                 bool find_unused_parameters, bool gradient_as_bucket_view, bool static_graph)
    : params_(std::move(params)),
      comm_(std::move(comm)),
      bucket_bytes_cap_(bucket_bytes_cap),
      first_bucket_bytes_cap_(first_bucket_bytes_cap),
      find_unused_(find_unused_parameters),
      grad_as_view_(gradient_as_bucket_view),
      static_graph_(static_graph) {
  TORCH_CHECK(!params_.empty(), "Reducer: no parameters require gradients");
  TORCH_CHECK(comm_ != nullptr, "Reducer: a communication backend is required");
  postscale_ = 1.0 / static_cast<double>(comm_->size());
  for (size_t i = 0; i < params_.size(); ++i) {
    const at::Tensor& p = params_[i];
    TORCH_CHECK(p.requires_grad(), "Reducer: parameter ", i, " does not require grad");
    TORCH_CHECK(!p.is_sparse(), "Reducer: sparse parameters are not supported");
    stats_.total_param_bytes += static_cast<int64_t>(p.nbytes());
  }
  build_buckets(bucket_indices);
  // One post-hook per parameter on its AccumulateGrad node: fires after .grad has been
  // written for this backward.
  for (size_t i = 0; i < params_.size(); ++i) {
    auto acc = torch::autograd::impl::grad_accumulator(params_[i]);
    TORCH_CHECK(acc, "Reducer: parameter ", i, " has no gradient accumulator (is it a leaf?)");
    uintptr_t key = acc->add_post_hook(std::make_unique<torch::autograd::utils::LambdaPostHook>(
        [this, i](const torch::autograd::variable_list& outputs, const torch::autograd::variable_list& /*inputs*/) {
          this->autograd_hook(i);
          return outputs;
        }));
    acc_to_index_[acc.get()] = i;
    hooks_.emplace_back(std::move(acc), key);
  }
  stats_.gradient_as_bucket_view = grad_as_view_;
  stats_.find_unused_parameters = find_unused_;
}

Reducer::~Reducer() {
  for (auto& h : hooks_) h.first->del_post_hook(h.second);
  hooks_.clear();
}

void Reducer::build_buckets(const std::vector<std::vector<int64_t>>& bucket_indices) {
  std::vector<char> seen(params_.size(), 0);
  std::vector<Bucket> fresh;
  fresh.reserve(bucket_indices.size());
  locs_.assign(params_.size(), Loc{0, 0});
  for (size_t b = 0; b < bucket_indices.size(); ++b) {
    const auto& idx = bucket_indices[b];
    TORCH_CHECK(!idx.empty(), "Reducer: empty bucket ", b);
    Bucket bk;
    const at::Tensor& first = params_.at(static_cast<size_t>(idx[0]));
    const int64_t esize = static_cast<int64_t>(first.element_size());
    const int64_t align_elems = std::max<int64_t>(1, kAlignBytes / esize);
    int64_t off = 0;
    for (int64_t gi : idx) {
      TORCH_CHECK(gi >= 0 && static_cast<size_t>(gi) < params_.size(), "Reducer: bad parameter index ", gi);
      TORCH_CHECK(!seen[static_cast<size_t>(gi)], "Reducer: parameter ", gi, " assigned to two buckets");
      seen[static_cast<size_t>(gi)] = 1;
      const at::Tensor& p = params_[static_cast<size_t>(gi)];
      TORCH_CHECK(p.scalar_type() == first.scalar_type() && p.device() == first.device(),
                  "Reducer: bucket ", b, " mixes dtypes or devices");
      off = (off + align_elems - 1) / align_elems * align_elems;
      bk.params.push_back(gi);
      bk.offsets.push_back(off);
      bk.lengths.push_back(p.numel());
      locs_[static_cast<size_t>(gi)] = Loc{b, bk.params.size() - 1};
      off += p.numel();
    }
    int64_t total = (off + align_elems - 1) / align_elems * align_elems;
    bk.flat = comm_->alloc_flat(total, first.scalar_type(), first.device());
    TORCH_CHECK(bk.flat.numel() >= total && bk.flat.is_contiguous(), "Reducer: backend returned a bad flat buffer");
    for (size_t s = 0; s < bk.params.size(); ++s) {
      const at::Tensor& p = params_[static_cast<size_t>(bk.params[s])];
      at::Tensor v;
      if (p.is_contiguous() || !p.is_non_overlapping_and_dense()) {
        v = bk.flat.narrow(0, bk.offsets[s], bk.lengths[s]).view(p.sizes());
      } else {
        // keep exotic-but-dense layouts (e.g. channels_last) so autograd's layout contract holds
        v = bk.flat.as_strided(p.sizes(), p.strides(), bk.flat.storage_offset() + bk.offsets[s]);
      }
      bk.views.push_back(std::move(v));
    }
    plan_chunks(bk);
    fresh.push_back(std::move(bk));
  }
  for (size_t i = 0; i < seen.size(); ++i) TORCH_CHECK(seen[i], "Reducer: parameter ", i, " is in no bucket");
  // carry over gradients that currently live in old bucket views
  if (!buckets_.empty()) {
    for (size_t i = 0; i < params_.size(); ++i) {
      at::Tensor& g = params_[i].mutable_grad();
      const Loc& nl = locs_[i];
      if (g.defined()) {
        fresh[nl.bucket].views[nl.slot].copy_(g);
        if (grad_as_view_) g = at::alias(fresh[nl.bucket].views[nl.slot]);
      }
    }
  }
  buckets_ = std::move(fresh);
  stats_.bucket_sizes_bytes.clear();
  stats_.bucket_indices.clear();
  for (auto& bk : buckets_) {
    int64_t bytes = 0;
    for (int64_t gi : bk.params) bytes += static_cast<int64_t>(params_[static_cast<size_t>(gi)].nbytes());
    stats_.bucket_sizes_bytes.push_back(bytes);
    stats_.bucket_indices.push_back(bk.params);
  }
}

void Reducer::plan_chunks(Bucket& bk) const {
  // Contiguous slot ranges in bucket order (= grad-ready order after the rebuild).  A chunk closes once it holds at
  // least `target` bytes; the remainder forms the last chunk.  Slot offsets are 16-byte aligned, so are the ranges.
  bk.chunks.clear();
  bk.slot_chunk.assign(bk.params.size(), 0);
  const int64_t esize = static_cast<int64_t>(bk.flat.element_size());
  const int64_t total = bk.flat.numel();
  const int64_t nmax = (has_comm_hook_ || max_chunks_ < 1) ? 1 : max_chunks_;
  const int64_t target = std::max<int64_t>(min_chunk_bytes_, (total * esize + nmax - 1) / nmax);
  Chunk cur;
  cur.first_slot = 0;
  cur.off = 0;
  for (size_t s = 0; s < bk.params.size(); ++s) {
    const int64_t end = (s + 1 < bk.params.size()) ? bk.offsets[s + 1] : total;
    const bool last_slot = (s + 1 == bk.params.size());
    const bool room = static_cast<int64_t>(bk.chunks.size()) + 1 < nmax;
    if (last_slot || (room && (end - cur.off) * esize >= target)) {
      cur.end_slot = s + 1;
      cur.len = end - cur.off;
      for (size_t q = cur.first_slot; q < cur.end_slot; ++q) bk.slot_chunk[q] = bk.chunks.size();
      bk.chunks.push_back(cur);
      cur = Chunk();
      cur.first_slot = s + 1;
      cur.off = end;
    }
  }
  // A last chunk much smaller than the target is not worth a kernel (and a cross-GPU barrier) of its own: the gradients that
  // arrive last are the ones nothing can hide, so they ride with the chunk before them.
  if (bk.chunks.size() >= 2 && bk.chunks.back().len * esize * 4 < target) {
    const Chunk tail = bk.chunks.back();
    bk.chunks.pop_back();
    Chunk& prev = bk.chunks.back();
    prev.end_slot = tail.end_slot;
    prev.len += tail.len;
    for (size_t q = tail.first_slot; q < tail.end_slot; ++q) bk.slot_chunk[q] = bk.chunks.size() - 1;
  }
}

void Reducer::set_chunking(int64_t min_chunk_bytes, int64_t max_chunks) {
  std::lock_guard<std::mutex> g(mu_);
  TORCH_CHECK(!expect_hooks_, "set_chunking must not run between forward and backward");
  min_chunk_bytes_ = std::max<int64_t>(min_chunk_bytes, 16);
  max_chunks_ = std::max<int64_t>(max_chunks, 1);
  for (auto& bk : buckets_) plan_chunks(bk);
}

void Reducer::set_fused_sgd(at::Tensor param_flat, at::Tensor momentum_flat, FusedSgd hyper, at::Tensor bcast, int64_t bcast_root) {
  std::lock_guard<std::mutex> g(mu_);
  TORCH_CHECK(!expect_hooks_, "set_fused_sgd must not run between forward and backward");
  if (!param_flat.defined()) {
    fused_param_ = at::Tensor();
    fused_momentum_ = at::Tensor();
    fused_bcast_ = at::Tensor();
    return;
  }
  TORCH_CHECK(buckets_.size() == 1 && !has_comm_hook_ && !find_unused_, "fused optimizer needs a single bucket, no comm hook, no unused-parameter search");
  TORCH_CHECK(param_flat.numel() == buckets_[0].flat.numel() && param_flat.scalar_type() == buckets_[0].flat.scalar_type() &&
                  param_flat.is_contiguous(), "fused optimizer: the flat parameter vector must mirror the bucket");
  TORCH_CHECK(hyper.momentum == 0 || (momentum_flat.defined() && momentum_flat.numel() == param_flat.numel()),
              "fused optimizer: momentum needs a flat buffer shaped like the bucket");
  fused_param_ = std::move(param_flat);
  fused_momentum_ = std::move(momentum_flat);
  fused_hyper_ = std::move(hyper);
  fused_bcast_ = std::move(bcast);
  fused_bcast_root_ = static_cast<int>(bcast_root);
}

void Reducer::resolve_timings() {
  // fold every iteration whose marks the device has passed into the statistics (never blocks)
  while (!unresolved_.empty()) {
    Marks& m = unresolved_.front();
    if (!(m.after_wait && m.after_wait->ready())) break;
    if (m.fwd_start && m.bwd_start) stats_.forward_us = m.bwd_start->us_since(*m.fwd_start);
    if (m.bwd_start && m.before_wait) stats_.backward_compute_us = m.before_wait->us_since(*m.bwd_start);
    if (m.first_ready && m.comm_end && m.comm_end->ready()) stats_.backward_comm_us = m.comm_end->us_since(*m.first_ready);
    if (m.before_wait) stats_.backward_comm_exposed_us = m.after_wait->us_since(*m.before_wait);
    if (m.bwd_start) stats_.backward_total_us = m.after_wait->us_since(*m.bwd_start);
    constexpr int64_t kStatsWarmup = 10;
    if (m.iteration > kStatsWarmup) {
      const double n = static_cast<double>(++stats_.timed_iterations);
      stats_.avg_forward_us += (stats_.forward_us - stats_.avg_forward_us) / n;
      stats_.avg_backward_compute_us += (stats_.backward_compute_us - stats_.avg_backward_compute_us) / n;
      stats_.avg_backward_comm_us += (stats_.backward_comm_us - stats_.avg_backward_comm_us) / n;
      stats_.avg_backward_comm_exposed_us += (stats_.backward_comm_exposed_us - stats_.avg_backward_comm_exposed_us) / n;
    }
    unresolved_.pop_front();
  }
  while (unresolved_.size() > 16) unresolved_.pop_front();
}

void Reducer::prepare_for_forward() {
  std::lock_guard<std::mutex> g(mu_);
  // A synchronised backward that started (some hook fired) but never reached its last bucket means some parameter
  // produced no gradient: the collective was not launched and ranks would silently diverge.  Same diagnosis and
  // advice as the reference's substrate (reducer "Expected to have finished reduction in the prior iteration…").
  if (expect_hooks_ && saw_first_hook_ && next_bucket_ < buckets_.size()) {
    std::string missing;
    int shown = 0;
    for (size_t i = 0; i < params_.size() && shown < 8; ++i)
      if (!ready_[i]) { missing += (shown++ ? ", " : "") + std::to_string(i); }
    expect_hooks_ = false;
    TORCH_CHECK(false,
                "Expected to have finished reduction in the prior iteration before starting a new one. This error indicates that "
                "your module has parameters that were not used in producing loss (parameter indices without a gradient: ",
                missing, shown == 8 ? ", …" : "",
                "). Enable unused parameter detection by passing `find_unused_parameters=True` to DistributedDataParallel, and make "
                "sure all `forward` outputs participate in calculating loss.");
  }
  ++stats_.num_iterations;
  timing_this_iter_ = !comm_->capturing();
  if (timing_this_iter_) resolve_timings();   // event queries are not allowed while a CUDA graph is being captured
  cur_ = Marks();
  cur_.iteration = stats_.num_iterations;
  if (timing_this_iter_) cur_.fwd_start = comm_->stamp();
}

void Reducer::prepare_for_backward(const std::vector<at::Tensor>& outputs) {
  std::lock_guard<std::mutex> g(mu_);
  expect_hooks_ = require_sync_;
  callback_queued_ = false;
  saw_first_hook_ = false;
  next_bucket_ = 0;
  ready_.assign(params_.size(), 0);
  locally_unused_.assign(params_.size(), 0);
  if (!ready_order_.empty()) prev_ready_order_ = ready_order_;
  ready_order_.clear();
  for (auto& bk : buckets_) {
    bk.pending = bk.params.size();
    bk.launched = false;
    bk.work.reset();
    bk.py_future = py::object();
    bk.hook_result = at::Tensor();
    bk.next_chunk = 0;
    for (auto& c : bk.chunks) {
      c.pending = c.end_slot - c.first_slot;
      c.work.reset();
    }
  }
  if (expect_hooks_ && find_unused_ && !outputs.empty()) search_unused(outputs);
}

void Reducer::search_unused(const std::vector<at::Tensor>& outputs) {
  // BFS over the autograd graph from the outputs; any parameter whose AccumulateGrad node is
  // unreachable will never fire its hook this iteration, so account for it now.
  std::unordered_set<torch::autograd::Node*> seen;
  std::deque<torch::autograd::Node*> queue;
  for (auto& o : outputs) {
    if (!o.defined()) continue;
    auto fn = o.grad_fn();
    if (fn && seen.insert(fn.get()).second) queue.push_back(fn.get());
  }
  std::vector<char> reachable(params_.size(), 0);
  while (!queue.empty()) {
    auto* n = queue.front();
    queue.pop_front();
    auto it = acc_to_index_.find(n);
    if (it != acc_to_index_.end()) reachable[it->second] = 1;
    for (const auto& e : n->next_edges()) {
      auto* nx = e.function.get();
      if (nx && seen.insert(nx).second) queue.push_back(nx);
    }
  }
  bool any = false;
  for (size_t i = 0; i < params_.size(); ++i) {
    if (reachable[i]) continue;
    locally_unused_[i] = 1;
    any = true;
  }
  if (!any) return;
  for (size_t i = 0; i < params_.size(); ++i) {
    if (!locally_unused_[i]) continue;
    const Loc& l = locs_[i];
    buckets_[l.bucket].views[l.slot].zero_();
    ready_[i] = 1;
    --buckets_[l.bucket].pending;
    --buckets_[l.bucket].chunks[buckets_[l.bucket].slot_chunk[l.slot]].pending;
  }
  // a bucket made entirely of unused params is launched by the first real hook (keeps the
  // collective inside backward and in bucket order)
}

void Reducer::autograd_hook(size_t index) {
  std::lock_guard<std::mutex> g(mu_);
  if (!expect_hooks_) return;  // no_sync(), or a backward that was not preceded by our forward
  if (!saw_first_hook_) {
    saw_first_hook_ = true;
    if (timing_this_iter_) cur_.bwd_start = comm_->stamp();
  }
  TORCH_CHECK(!ready_[index] || static_graph_,
              "pdt Reducer: parameter ", index,
              " was marked ready twice in one backward. The module was probably run more than once per "
              "iteration, or the same parameters are reused outside DistributedDataParallel.forward.");
  if (ready_[index]) return;
  mark_variable_ready(index);
}

void Reducer::mark_variable_ready(size_t index) {
  ready_[index] = 1;
  ready_order_.push_back(static_cast<int64_t>(index));
  const Loc& l = locs_[index];
  Bucket& bk = buckets_[l.bucket];
  at::Tensor& grad = params_[index].mutable_grad();
  at::Tensor& view = bk.views[l.slot];
  if (grad.defined()) {
    if (!same_view(grad, view)) {
      view.copy_(grad);
      ++stats_.copies_into_bucket;
      if (grad_as_view_) grad = at::alias(view);
    }
  } else {
    view.zero_();
    locally_unused_[index] = 1;
  }
  TORCH_CHECK(bk.pending > 0, "pdt Reducer: bucket bookkeeping underflow");
  --bk.pending;
  --bk.chunks[bk.slot_chunk[l.slot]].pending;
  launch_ready_buckets();
}

void Reducer::launch_ready_buckets() {
  // strictly in (bucket, chunk) order — the same order on every rank
  while (next_bucket_ < buckets_.size()) {
    Bucket& bk = buckets_[next_bucket_];
    const bool whole = has_comm_hook_ || defer_comm_;  // a hook / the optimizer consumes whole buckets
    if (whole) {
      if (bk.pending != 0) break;
      launch_bucket(next_bucket_);
    } else {
      while (bk.next_chunk < bk.chunks.size() && bk.chunks[bk.next_chunk].pending == 0) {
        launch_chunk(next_bucket_, bk.next_chunk);
        ++bk.next_chunk;
      }
      if (bk.next_chunk < bk.chunks.size()) break;
      bk.launched = true;
      ++stats_.num_buckets_reduced;
    }
    ++next_bucket_;
  }
  if (next_bucket_ == buckets_.size() && !callback_queued_) {
    callback_queued_ = true;
    torch::autograd::Engine::get_default_engine().queue_callback([this] { this->finalize_backward(); });
  }
}

void Reducer::launch_chunk(size_t b, size_t c) {
  Bucket& bk = buckets_[b];
  Chunk& ch = bk.chunks[c];
  if (timing_this_iter_ && !cur_.first_ready) cur_.first_ready = comm_->stamp();
  at::Tensor piece = (ch.off == 0 && ch.len == bk.flat.numel()) ? bk.flat : bk.flat.narrow(0, ch.off, ch.len);
  if (fused_param_.defined()) {
    const bool last = (b + 1 == buckets_.size()) && (c + 1 == bk.chunks.size());
    at::Tensor mom = fused_momentum_.defined() ? fused_momentum_.narrow(0, ch.off, ch.len) : at::Tensor();
    ch.work = comm_->allreduce_sgd(piece, fused_param_.narrow(0, ch.off, ch.len), mom, fused_hyper_,
                                   last ? fused_bcast_ : at::Tensor(), fused_bcast_root_);
  } else {
    ch.work = comm_->allreduce(piece, ReduceOp::SUM, postscale_);
  }
  if (timing_this_iter_) cur_.comm_end = comm_->stamp(/*on_comm_stream=*/true);
}

void Reducer::launch_bucket(size_t b) {
  Bucket& bk = buckets_[b];
  if (timing_this_iter_ && !cur_.first_ready) cur_.first_ready = comm_->stamp();
  bk.launched = true;
  ++stats_.num_buckets_reduced;
  if (has_comm_hook_) {
    py::gil_scoped_acquire gil;
    GradBucket gb;
    gb.index = static_cast<int64_t>(b);
    gb.is_last = (b + 1 == buckets_.size());
    gb.buffer = bk.flat;
    gb.gradients = bk.views;
    for (int64_t gi : bk.params) gb.parameters.push_back(params_[static_cast<size_t>(gi)]);
    gb.offsets = bk.offsets;
    gb.lengths = bk.lengths;
    bk.py_future = comm_hook_(py::cast(std::move(gb)));
  } else if (!defer_comm_) {
    bk.work = comm_->allreduce(bk.flat, ReduceOp::SUM, postscale_);
  }
  if (timing_this_iter_) cur_.comm_end = comm_->stamp(/*on_comm_stream=*/true);
}

void Reducer::finalize_backward() {
  std::lock_guard<std::mutex> g(mu_);
  if (!expect_hooks_) return;
  expect_hooks_ = false;
  if (timing_this_iter_) cur_.before_wait = comm_->stamp();
  int64_t launched_chunks = 0;
  for (size_t b = 0; b < buckets_.size(); ++b) {
    Bucket& bk = buckets_[b];
    TORCH_CHECK(bk.launched, "pdt Reducer: bucket ", b, " was never launched");
    if (has_comm_hook_) {
      py::gil_scoped_acquire gil;
      py::object res = bk.py_future;
      if (!THPVariable_Check(res.ptr()) && py::hasattr(res, "wait")) res = res.attr("wait")();
      TORCH_CHECK(THPVariable_Check(res.ptr()), "comm hook must return a Tensor or a future whose wait() returns a Tensor");
      at::Tensor out = res.cast<at::Tensor>();
      TORCH_CHECK(out.numel() == bk.flat.numel(), "comm hook must return a tensor shaped like bucket.buffer");
      if (out.data_ptr() != bk.flat.data_ptr()) bk.flat.copy_(out.view_as(bk.flat));
      bk.py_future = py::object();
    } else if (bk.work) {
      bk.work->wait();
      bk.work.reset();
      ++launched_chunks;
    }
    for (auto& c : bk.chunks)
      if (c.work) {
        c.work->wait();   // CUDA: the compute stream waits for the comm stream; never blocks the host
        c.work.reset();
        ++launched_chunks;
      }
  }
  stats_.reduce_chunks = launched_chunks;
  if (timing_this_iter_) {
    cur_.after_wait = comm_->stamp();
    unresolved_.push_back(cur_);
  }
  // globally-unused detection: a parameter unused on every rank keeps grad=None
  std::vector<char> globally_unused(params_.size(), 0);
  bool any_local_unused = false;
  for (char c : locally_unused_) any_local_unused |= (c != 0);
  if (find_unused_) {
    at::Tensor used = at::empty({static_cast<int64_t>(params_.size())}, at::kInt);
    auto* up = used.data_ptr<int32_t>();
    for (size_t i = 0; i < params_.size(); ++i) up[i] = locally_unused_[i] ? 0 : 1;
    at::Tensor dev = comm_->is_cuda() ? used.to(params_[0].device()) : used;
    auto w = comm_->allreduce(dev, ReduceOp::SUM, 1.0);
    w->wait();
    at::Tensor host = dev.to(at::kCPU);
    auto* hp = host.data_ptr<int32_t>();
    for (size_t i = 0; i < params_.size(); ++i) globally_unused[i] = (hp[i] == 0);
  } else {
    (void)any_local_unused;
  }
  for (size_t i = 0; i < params_.size(); ++i) {
    const Loc& l = locs_[i];
    at::Tensor& view = buckets_[l.bucket].views[l.slot];
    at::Tensor& grad = params_[i].mutable_grad();
    if (globally_unused[i]) continue;  // leave whatever the user had (normally None)
    if (grad_as_view_) {
      if (!same_view(grad, view)) grad = at::alias(view);
    } else {
      if (!grad.defined()) grad = at::empty_like(params_[i]);
      if (!same_view(grad, view)) grad.copy_(view);
    }
  }
  stats_.grad_ready_order = ready_order_;
}

bool Reducer::should_rebuild() const {
  std::lock_guard<std::mutex> g(mu_);
  if (has_rebuilt_ || find_unused_) return false;  // with unused params the order is not stable
  const auto& order = ready_order_.empty() ? prev_ready_order_ : ready_order_;
  return order.size() == params_.size();
}

std::vector<std::vector<int64_t>> Reducer::propose_rebuild() const {
  std::lock_guard<std::mutex> g(mu_);
  const auto& order = ready_order_.empty() ? prev_ready_order_ : ready_order_;
  TORCH_CHECK(order.size() == params_.size(), "propose_rebuild: no complete grad-ready order recorded yet");
  std::vector<PlanInput> in;
  in.reserve(params_.size());
  for (auto& p : params_) in.push_back(PlanInput{static_cast<int64_t>(p.nbytes()), group_key(p)});
  return plan_buckets(in, {first_bucket_bytes_cap_, bucket_bytes_cap_}, order).buckets;
}

void Reducer::apply_rebuild(const std::vector<std::vector<int64_t>>& bucket_indices) {
  std::lock_guard<std::mutex> g(mu_);
  TORCH_CHECK(!expect_hooks_, "apply_rebuild must not run between forward and backward");
  bool same = bucket_indices.size() == buckets_.size();
  for (size_t b = 0; same && b < buckets_.size(); ++b) same = (buckets_[b].params == bucket_indices[b]);
  if (!same) build_buckets(bucket_indices);
  has_rebuilt_ = true;
  stats_.has_rebuilt = true;
  ++stats_.num_rebuilds;
}

void Reducer::register_comm_hook(py::object hook) {
  std::lock_guard<std::mutex> g(mu_);
  TORCH_CHECK(!has_comm_hook_, "register_comm_hook can only be called once");
  comm_hook_ = std::move(hook);
  has_comm_hook_ = true;
}

ReducerStats Reducer::stats() const {
  std::lock_guard<std::mutex> g(mu_);
  if (!comm_->capturing()) const_cast<Reducer*>(this)->resolve_timings();
  ReducerStats s = stats_;
  if (s.grad_ready_order.empty()) s.grad_ready_order = prev_ready_order_;
  return s;
}

std::vector<at::Tensor> Reducer::bucket_buffers() const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<at::Tensor> out;
  for (auto& b : buckets_) out.push_back(b.flat);
  return out;
}

std::vector<at::Tensor> Reducer::grad_views() const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<at::Tensor> out;
  out.reserve(params_.size());
  for (size_t i = 0; i < params_.size(); ++i) out.push_back(buckets_[locs_[i].bucket].views[locs_[i].slot]);
  return out;
}

bool Reducer::grads_are_views() const {
  std::lock_guard<std::mutex> g(mu_);
  for (size_t i = 0; i < params_.size(); ++i) {
    const Loc& l = locs_[i];
    if (!same_view(params_[i].grad(), buckets_[l.bucket].views[l.slot])) return false;
  }
  return true;
}

void Reducer::install_grad_views(bool zero) {
  std::lock_guard<std::mutex> g(mu_);
  if (zero) for (auto& b : buckets_) b.flat.zero_();
  for (size_t i = 0; i < params_.size(); ++i) {
    const Loc& l = locs_[i];
    at::Tensor& grad = params_[i].mutable_grad();
    const at::Tensor& view = buckets_[l.bucket].views[l.slot];
    if (!same_view(grad, view)) {
      if (grad.defined() && !zero) buckets_[l.bucket].views[l.slot].copy_(grad);
      grad = at::alias(view);
    }
  }
}

}  // namespace pdt
