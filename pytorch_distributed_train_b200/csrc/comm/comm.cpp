#include "comm.h"

#include <c10/core/GradMode.h>
#include <c10/util/Exception.h>

#include <chrono>

namespace pdt {

DType to_dtype(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return DType::F32;
    case at::kDouble: return DType::F64;
    case at::kHalf: return DType::F16;
    case at::kBFloat16: return DType::BF16;
    case at::kChar: return DType::I8;
    case at::kByte: return DType::U8;
    case at::kShort: return DType::I16;
    case at::kInt: return DType::I32;
    case at::kLong: return DType::I64;
    case at::kBool: return DType::BOOL;
    default: TORCH_CHECK(false, "pdt: unsupported dtype for collectives: ", c10::toString(t));
  }
}

static const char* kUnsupported = " is not implemented by this backend";
std::shared_ptr<CommWork> Comm::reduce(at::Tensor, ReduceOp, int) { TORCH_CHECK(false, "reduce", kUnsupported); }
std::shared_ptr<CommWork> Comm::reduce_scatter(at::Tensor, at::Tensor, ReduceOp) { TORCH_CHECK(false, "reduce_scatter", kUnsupported); }
std::shared_ptr<CommWork> Comm::gather(at::Tensor, at::Tensor, int) { TORCH_CHECK(false, "gather", kUnsupported); }
std::shared_ptr<CommWork> Comm::scatter(at::Tensor, at::Tensor, int) { TORCH_CHECK(false, "scatter", kUnsupported); }
std::shared_ptr<CommWork> Comm::alltoall(at::Tensor, at::Tensor) { TORCH_CHECK(false, "alltoall", kUnsupported); }
std::shared_ptr<CommWork> Comm::send(at::Tensor, int) { TORCH_CHECK(false, "send", kUnsupported); }
std::shared_ptr<CommWork> Comm::recv(at::Tensor, int) { TORCH_CHECK(false, "recv", kUnsupported); }

namespace {
class HostStamp : public DeviceStamp {
 public:
  HostStamp() : t_(std::chrono::steady_clock::now()) {}
  bool ready() override { return true; }
  double us_since(DeviceStamp& earlier) override {
    return std::chrono::duration<double, std::micro>(t_ - static_cast<HostStamp&>(earlier).t_).count();
  }

 private:
  std::chrono::steady_clock::time_point t_;
};

class DoneWork : public CommWork {
 public:
  void wait() override {}
  void synchronize() override {}
  bool is_completed() override { return true; }
};
}  // namespace

std::shared_ptr<DeviceStamp> Comm::stamp(bool) { return std::make_shared<HostStamp>(); }

std::shared_ptr<CommWork> Comm::allreduce_sgd(at::Tensor grad, at::Tensor param, at::Tensor momentum_buf, const FusedSgd& h,
                                              at::Tensor bcast, int bcast_root) {
  // reference composition (CPU backend, NCCL baseline): three steps where SymmComm needs one kernel
  allreduce(grad, ReduceOp::SUM, 1.0 / static_cast<double>(size()))->wait();
  c10::NoGradGuard ng;
  at::Tensor g = h.weight_decay != 0 ? grad.add(param, h.weight_decay) : grad;
  if (h.momentum != 0) {
    TORCH_CHECK(momentum_buf.defined() && momentum_buf.numel() == grad.numel(), "allreduce_sgd: momentum buffer required");
    if (h.first_step) momentum_buf.copy_(g);
    else momentum_buf.mul_(h.momentum).add_(g, 1.0 - h.dampening);
    g = h.nesterov ? g.add(momentum_buf, h.momentum) : momentum_buf;
  }
  if (h.lr_tensor.defined()) param.sub_(g * h.lr_tensor);
  else param.add_(g, -h.lr);
  if (bcast.defined() && bcast.numel() > 0 && size() > 1) broadcast(bcast, bcast_root)->wait();
  return std::make_shared<DoneWork>();
}

void Comm::record(const char* op, const at::Tensor* t) {
  std::lock_guard<std::mutex> g(rec_mu_);
  Record r;
  r.seq = seq_++;
  r.op = op;
  r.numel = t ? t->numel() : 0;
  r.dtype = t ? c10::toString(t->scalar_type()) : "-";
  r.t_enqueue = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
  if (ring_.size() < kRing) ring_.push_back(std::move(r));
  else ring_[r.seq % kRing] = std::move(r);
}

std::vector<Comm::Record> Comm::flight_records() const {
  std::lock_guard<std::mutex> g(rec_mu_);
  std::vector<Record> out = ring_;
  std::sort(out.begin(), out.end(), [](const Record& a, const Record& b) { return a.seq < b.seq; });
  return out;
}

// ---------------------------------------------------------------------------------------
namespace {

// Keeps the tensors alive until the worker is done with their memory.
class CpuWork : public CommWork {
 public:
  CpuWork(std::shared_ptr<Work> w, std::vector<at::Tensor> keep, Millis timeout,
          std::function<void()> epilogue = nullptr)
      : w_(std::move(w)), keep_(std::move(keep)), timeout_(timeout), epilogue_(std::move(epilogue)) {}
  void wait() override {
    w_->wait(timeout_);
    if (epilogue_ && !ran_) { ran_ = true; epilogue_(); }
    keep_.clear();
  }
  void synchronize() override { wait(); }
  bool is_completed() override { return w_->is_completed(); }

 private:
  std::shared_ptr<Work> w_;
  std::vector<at::Tensor> keep_;
  Millis timeout_;
  std::function<void()> epilogue_;
  bool ran_ = false;
};

void check_cpu(const at::Tensor& t, const char* what) {
  TORCH_CHECK(t.device().is_cpu(), "pdt cpu backend: ", what, " tensor must live on the CPU (got ", t.device(), ")");
  TORCH_CHECK(t.is_contiguous(), "pdt cpu backend: ", what, " tensor must be contiguous");
}

}  // namespace

std::shared_ptr<CommWork> CpuComm::allreduce(at::Tensor t, ReduceOp op, double postscale) {
  check_cpu(t, "allreduce");
  record("allreduce", &t);
  auto w = be_->allreduce(t.data_ptr(), static_cast<size_t>(t.numel()), to_dtype(t.scalar_type()), op);
  std::function<void()> epi;
  if (postscale != 1.0) epi = [t, postscale]() mutable { t.mul_(postscale); };
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{t}, be_->timeout() + Millis(1000), epi);
}
std::shared_ptr<CommWork> CpuComm::broadcast(at::Tensor t, int root) {
  check_cpu(t, "broadcast");
  record("broadcast", &t);
  auto w = be_->broadcast(t.data_ptr(), t.nbytes(), root);
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{t}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::allgather(at::Tensor out, at::Tensor in) {
  check_cpu(out, "allgather output");
  check_cpu(in, "allgather input");
  TORCH_CHECK(out.numel() == in.numel() * size() && out.scalar_type() == in.scalar_type(),
              "allgather: output must hold world_size × input elements of the same dtype");
  record("allgather", &in);
  auto w = be_->allgather(in.data_ptr(), out.data_ptr(), in.nbytes());
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{out, in}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::reduce(at::Tensor t, ReduceOp op, int root) {
  check_cpu(t, "reduce");
  record("reduce", &t);
  auto w = be_->reduce(t.data_ptr(), static_cast<size_t>(t.numel()), to_dtype(t.scalar_type()), op, root);
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{t}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::reduce_scatter(at::Tensor out, at::Tensor in, ReduceOp op) {
  check_cpu(out, "reduce_scatter output");
  check_cpu(in, "reduce_scatter input");
  TORCH_CHECK(in.numel() == out.numel() * size(), "reduce_scatter: input must hold world_size × output elements");
  record("reduce_scatter", &in);
  auto w = be_->reduce_scatter(in.data_ptr(), out.data_ptr(), static_cast<size_t>(out.numel()), to_dtype(in.scalar_type()), op);
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{out, in}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::gather(at::Tensor out, at::Tensor in, int root) {
  check_cpu(in, "gather input");
  if (rank() == root) {
    check_cpu(out, "gather output");
    TORCH_CHECK(out.numel() == in.numel() * size(), "gather: output must hold world_size × input elements");
  }
  record("gather", &in);
  auto w = be_->gather(in.data_ptr(), rank() == root ? out.data_ptr() : nullptr, in.nbytes(), root);
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{out, in}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::scatter(at::Tensor out, at::Tensor in, int root) {
  check_cpu(out, "scatter output");
  if (rank() == root) {
    check_cpu(in, "scatter input");
    TORCH_CHECK(in.numel() == out.numel() * size(), "scatter: input must hold world_size × output elements");
  }
  record("scatter", &out);
  auto w = be_->scatter(rank() == root ? in.data_ptr() : nullptr, out.data_ptr(), out.nbytes(), root);
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{out, in}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::alltoall(at::Tensor out, at::Tensor in) {
  check_cpu(out, "alltoall output");
  check_cpu(in, "alltoall input");
  TORCH_CHECK(in.numel() == out.numel() && in.numel() % size() == 0, "alltoall: equal splits required");
  record("alltoall", &in);
  auto w = be_->alltoall(in.data_ptr(), out.data_ptr(), in.nbytes() / size());
  return std::make_shared<CpuWork>(w, std::vector<at::Tensor>{out, in}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::send(at::Tensor t, int dst) {
  check_cpu(t, "send");
  record("send", &t);
  return std::make_shared<CpuWork>(be_->send(t.data_ptr(), t.nbytes(), dst), std::vector<at::Tensor>{t}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::recv(at::Tensor t, int src) {
  check_cpu(t, "recv");
  record("recv", &t);
  return std::make_shared<CpuWork>(be_->recv(t.data_ptr(), t.nbytes(), src), std::vector<at::Tensor>{t}, be_->timeout() + Millis(1000));
}
std::shared_ptr<CommWork> CpuComm::barrier() {
  record("barrier", nullptr);
  return std::make_shared<CpuWork>(be_->barrier(), std::vector<at::Tensor>{}, be_->timeout() + Millis(1000));
}

}  // namespace pdt
