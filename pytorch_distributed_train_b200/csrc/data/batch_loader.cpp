#include "batch_loader.h"

#include <c10/util/Exception.h>

#include <cstring>

#if PDT_WITH_CUDA
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAEvent.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#endif

namespace pdt {

namespace {
// uint8 → float32 × scale (ToTensor).  78,400 elements per MNIST batch: the scalar loop is ~0.2 ms, the AVX2 one ~10 µs —
// the difference between a loader that keeps up with a 0.1 ms training step and one that does not.
#if defined(__x86_64__)
__attribute__((target("avx2"), optimize("O3"))) void u8_to_f32_avx2(const uint8_t* __restrict p, float* __restrict o, int64_t n, float sc) {
  for (int64_t k = 0; k < n; ++k) o[k] = static_cast<float>(p[k]) * sc;
}
#endif
__attribute__((optimize("O3"))) void u8_to_f32_generic(const uint8_t* __restrict p, float* __restrict o, int64_t n, float sc) {
  for (int64_t k = 0; k < n; ++k) o[k] = static_cast<float>(p[k]) * sc;
}
void u8_to_f32(const uint8_t* p, float* o, int64_t n, float sc) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return u8_to_f32_avx2(p, o, n, sc);
#endif
  u8_to_f32_generic(p, o, n, sc);
}
}  // namespace

BatchStager::BatchStager(at::Tensor data, at::Tensor targets, std::vector<int64_t> sample_shape, int64_t batch_size, bool drop_last,
                         double scale, int64_t depth, bool pin_memory, int device)
    : data_(std::move(data)), targets_(std::move(targets)), sample_shape_(std::move(sample_shape)), batch_(batch_size),
      depth_(std::max<int64_t>(depth, 2)), drop_last_(drop_last), pin_(pin_memory), scale_(scale), device_(device) {
  TORCH_CHECK(data_.device().is_cpu() && targets_.device().is_cpu() && data_.is_contiguous() && targets_.is_contiguous(),
              "BatchStager: dataset tensors must be contiguous CPU tensors");
  TORCH_CHECK(data_.scalar_type() == at::kByte || data_.scalar_type() == at::kFloat, "BatchStager: data must be uint8 or float32");
  TORCH_CHECK(data_.dim() >= 1 && targets_.dim() == 1 && data_.size(0) == targets_.size(0), "BatchStager: data/targets disagree");
  TORCH_CHECK(batch_ > 0, "BatchStager: batch size must be positive");
  row_elems_ = data_.numel() / std::max<int64_t>(data_.size(0), 1);
  int64_t want = 1;
  for (auto d : sample_shape_) want *= d;
  TORCH_CHECK(want == row_elems_, "BatchStager: sample_shape does not match the dataset rows");
  ring_.resize(static_cast<size_t>(depth_));
  for (auto& s : ring_) alloc_slot(s);
}

BatchStager::~BatchStager() { stop_worker(); }

void BatchStager::alloc_slot(Slot& s) {
  std::vector<int64_t> shp{batch_};
  shp.insert(shp.end(), sample_shape_.begin(), sample_shape_.end());
  auto io = at::TensorOptions().dtype(at::kFloat).device(at::kCPU).pinned_memory(pin_);
  auto to = at::TensorOptions().dtype(targets_.scalar_type()).device(at::kCPU).pinned_memory(pin_);
  s.images = at::empty(shp, io);
  s.targets = at::empty({batch_}, to);
}

void BatchStager::stop_worker() {
  {
    std::lock_guard<std::mutex> g(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  if (th_.joinable()) th_.join();
  stop_ = false;
  running_ = false;
}

void BatchStager::start(at::Tensor indices) {
  TORCH_CHECK(indices.device().is_cpu() && indices.scalar_type() == at::kLong && indices.dim() == 1, "BatchStager.start: int64 CPU index vector expected");
  stop_worker();
  indices_ = indices.contiguous();
  const int64_t n = indices_.numel();
  nbatches_ = drop_last_ ? n / batch_ : (n + batch_ - 1) / batch_;
  produce_ = consume_ = 0;
  last_handed_ = -1;
  ++epoch_;
  for (auto& s : ring_) {
    if (s.state == 2) s.state = 0;   // a batch of the abandoned epoch may still be in user hands: the retention check covers it
    if (s.state == 1) s.state = 0;
  }
  running_ = true;
  th_ = std::thread([this] { this->worker(); });
}

void BatchStager::worker() {
#if PDT_WITH_CUDA
  if (pin_ && device_ >= 0) cudaSetDevice(device_);
#endif
  const int64_t n = indices_.numel();
  const int64_t* idx = indices_.data_ptr<int64_t>();
  const int64_t N = data_.size(0);
  const size_t tsize = targets_.element_size();
  for (int64_t b = 0; b < nbatches_; ++b) {
    Slot& s = ring_[static_cast<size_t>(b % depth_)];
    std::shared_ptr<void> ev;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || s.state == 0; });
      if (stop_) return;
      ev = std::move(s.event);
      s.event.reset();
    }
#if PDT_WITH_CUDA
    if (ev) static_cast<at::cuda::CUDAEvent*>(ev.get())->synchronize();   // the consumer's copies out of this slot are done
#endif
    // a batch somebody still holds (list(loader), a stashed view) is never overwritten: the slot gets fresh buffers
    if (s.images.use_count() > 1 || s.targets.use_count() > 1 || s.images.storage().use_count() > 1 || s.targets.storage().use_count() > 1)
      alloc_slot(s);
    const int64_t lo = b * batch_, hi = std::min(n, lo + batch_);
    float* out = s.images.data_ptr<float>();
    char* tout = static_cast<char*>(s.targets.data_ptr());
    const char* tin = static_cast<const char*>(targets_.data_ptr());
    if (data_.scalar_type() == at::kByte) {
      const uint8_t* src = data_.data_ptr<uint8_t>();
      const float sc = static_cast<float>(scale_);
      for (int64_t i = lo; i < hi; ++i) {
        const int64_t r = idx[i];
        TORCH_CHECK(r >= 0 && r < N, "BatchStager: index ", r, " out of range");
        const uint8_t* p = src + r * row_elems_;
        float* o = out + (i - lo) * row_elems_;
        u8_to_f32(p, o, row_elems_, sc);
        std::memcpy(tout + (i - lo) * tsize, tin + r * tsize, tsize);
      }
    } else {
      const float* src = data_.data_ptr<float>();
      const float sc = static_cast<float>(scale_);
      for (int64_t i = lo; i < hi; ++i) {
        const int64_t r = idx[i];
        TORCH_CHECK(r >= 0 && r < N, "BatchStager: index ", r, " out of range");
        const float* p = src + r * row_elems_;
        float* o = out + (i - lo) * row_elems_;
        if (sc == 1.f) std::memcpy(o, p, static_cast<size_t>(row_elems_) * sizeof(float));
        else for (int64_t k = 0; k < row_elems_; ++k) o[k] = p[k] * sc;
        std::memcpy(tout + (i - lo) * tsize, tin + r * tsize, tsize);
      }
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      s.rows = hi - lo;
      s.state = 1;
      ++produce_;
    }
    cv_.notify_all();
  }
}

bool BatchStager::next(at::Tensor* images, at::Tensor* targets) {
  // the consumer moves on: everything it enqueued on its current stream for the previous batch precedes this event
  if (last_handed_ >= 0) {
    Slot& p = ring_[static_cast<size_t>(last_handed_ % depth_)];
    std::shared_ptr<void> ev;
#if PDT_WITH_CUDA
    if (pin_ && device_ >= 0) {
      c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device_));
      auto* e = new at::cuda::CUDAEvent(cudaEventDisableTiming);
      e->record(c10::cuda::getCurrentCUDAStream(static_cast<c10::DeviceIndex>(device_)));
      ev = std::shared_ptr<void>(e, [](void* q) { delete static_cast<at::cuda::CUDAEvent*>(q); });
    }
#endif
    {
      std::lock_guard<std::mutex> g(mu_);
      p.event = std::move(ev);
      p.state = 0;
    }
    cv_.notify_all();
    last_handed_ = -1;
  }
  if (consume_ >= nbatches_) return false;
  Slot& s = ring_[static_cast<size_t>(consume_ % depth_)];
  {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return s.state == 1; });
    s.state = 2;
  }
  *images = s.rows == batch_ ? s.images : s.images.narrow(0, 0, s.rows);
  *targets = s.rows == batch_ ? s.targets : s.targets.narrow(0, 0, s.rows);
  last_handed_ = consume_;
  ++consume_;
  return true;
}

}  // namespace pdt
