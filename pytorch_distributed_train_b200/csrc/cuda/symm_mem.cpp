#include "symm_mem.h"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <random>

#include "../common/net.h"
#include "cuda_utils.h"

namespace pdt {

namespace {
size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

std::string hex_token() {
  std::random_device rd;
  char buf[32];
  snprintf(buf, sizeof(buf), "%08x%08x", rd(), rd());
  return buf;
}
}  // namespace

SymmetricHeap::SymmetricHeap(std::shared_ptr<Store> store, int rank, int world, int device, size_t heap_bytes, Millis timeout)
    : rank_(rank), world_(world), device_(device) {
  if (world < 1 || world > kSymmMaxWorld)
    throw std::invalid_argument("SymmetricHeap: world size must be in [1, " + std::to_string(kSymmMaxWorld) + "] (one NVSwitch domain)");
  PDT_CUDA_CHECK(cudaSetDevice(device));
  PDT_CUDA_CHECK(cudaFree(nullptr));  // make sure the primary context exists
  // ---- layout --------------------------------------------------------------------------------
  signal_bytes_ = round_up(static_cast<size_t>(kSymmChannels) * kSymmMaxBlocks * kSymmMaxWorld * sizeof(uint32_t), 4096);
  size_t off = signal_bytes_;
  const size_t staging_cfg[kSymmChannels] = {16u << 20, 1u << 20, 1u << 20, 16u << 20};
  for (int c = 0; c < kSymmChannels; ++c) {
    staging_off_[c] = off;
    staging_half_[c] = staging_cfg[c];
    off += 2 * staging_cfg[c];
  }
  user_off_ = round_up(off, 1 << 16);
  if (heap_bytes < user_off_ + (1u << 20)) heap_bytes = user_off_ + (64u << 20);
  heap_bytes_ = heap_bytes;
  peer_base_.assign(world, nullptr);

  // ---- map every rank's heap -------------------------------------------------------------------
  const char* force_ipc = getenv("PDT_SYMM_FORCE_IPC");
  bool done = false;
  if (!(force_ipc && force_ipc[0] == '1')) {
    try {
      setup_vmm(store, timeout);
      done = true;
    } catch (const std::exception& e) {
      // every rank takes the same branch: agree through the store
      store->set("symm/vmm_fail/" + std::to_string(rank_), e.what());
    }
    store->set("symm/vmm_done/" + std::to_string(rank_), done ? "1" : "0");
    bool all = true;
    for (int r = 0; r < world_; ++r) all &= (store->get("symm/vmm_done/" + std::to_string(r)) == "1");
    if (!all) {
      if (done) throw std::runtime_error("SymmetricHeap: VMM mapping succeeded locally but failed on a peer rank");
      done = false;
    }
  }
  if (!done) setup_ipc(store, timeout);

  // ---- local bookkeeping -----------------------------------------------------------------------
  PDT_CUDA_CHECK(cudaMemset(peer_base_[rank_], 0, user_off_));
  PDT_CUDA_CHECK(cudaMalloc(&epochs_, sizeof(uint32_t) * kSymmChannels * kSymmMaxBlocks));
  PDT_CUDA_CHECK(cudaMemset(epochs_, 0, sizeof(uint32_t) * kSymmChannels * kSymmMaxBlocks));
  PDT_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&status_host_), sizeof(int), cudaHostAllocMapped));
  *status_host_ = 0;
  PDT_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&status_dev_), status_host_, 0));
  PDT_CUDA_CHECK(cudaDeviceSynchronize());
  free_[user_off_] = heap_bytes_ - user_off_;
  if (const char* t = getenv("PDT_SYMM_TIMEOUT_S")) timeout_ns_ = static_cast<unsigned long long>(atof(t) * 1e9);
  // nobody may touch a peer's pad before that peer has zeroed it
  store_barrier(*store, "symm/ready", rank_, world_, timeout);
}

SymmetricHeap::~SymmetricHeap() {
  if (process_exiting()) return;  // interpreter shutdown: the driver reclaims everything
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  if (epochs_) cudaFree(epochs_);
  if (status_host_) cudaFreeHost(status_host_);
  if (vmm_) {
    try {
      const DriverApi& d = driver();
      if (mc_base_) {
        d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(mc_base_), mc_size_);
        d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(mc_base_), mc_size_);
        d.cuMemRelease(mc_handle_);
      }
      for (int r = 0; r < world_; ++r) {
        if (!peer_base_[r]) continue;
        d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(peer_base_[r]), heap_bytes_);
        d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(peer_base_[r]), heap_bytes_);
        if (r < static_cast<int>(handles_.size()) && handles_[r]) d.cuMemRelease(handles_[r]);
      }
    } catch (...) {
    }
  } else {
    for (int r = 0; r < world_; ++r) {
      if (!peer_base_[r]) continue;
      if (r == rank_) cudaFree(peer_base_[r]);
      else cudaIpcCloseMemHandle(peer_base_[r]);
    }
  }
}

// Pairwise fd swap: the higher rank dials the lower one; both directions go over that one
// connection. Socket names live in the abstract AF_UNIX namespace and are unique per job.
void SymmetricHeap::exchange_fds(std::shared_ptr<Store> store, const std::string& tag, int my_fd,
                                 std::vector<int>* peer_fds, Millis timeout) {
  peer_fds->assign(world_, -1);
  if (rank_ == 0) store->set("symm/token/" + tag, hex_token());
  std::string token = store->get("symm/token/" + tag);
  auto name = [&](int r) { return "pdt-symm-" + token + "-" + tag + "-" + std::to_string(r); };
  Fd lfd = unix_listen(name(rank_));
  store->set("symm/listening/" + tag + "/" + std::to_string(rank_), "1");
  for (int p = 0; p < rank_; ++p) {
    store->get("symm/listening/" + tag + "/" + std::to_string(p));
    Fd s = unix_connect(name(p), timeout);
    int32_t me = rank_;
    send_all(s.get(), &me, 4, timeout);
    send_fd(s.get(), my_fd, timeout);
    (*peer_fds)[p] = recv_fd(s.get(), timeout);
  }
  for (int k = rank_ + 1; k < world_; ++k) {
    Fd s = tcp_accept(lfd.get(), timeout);
    int32_t who = -1;
    recv_all(s.get(), &who, 4, timeout);
    if (who <= rank_ || who >= world_) throw std::runtime_error("symm fd exchange: unexpected peer " + std::to_string(who));
    (*peer_fds)[who] = recv_fd(s.get(), timeout);
    send_fd(s.get(), my_fd, timeout);
  }
}

void SymmetricHeap::setup_vmm(std::shared_ptr<Store> store, Millis timeout) {
  const DriverApi& d = driver();
  CUdevice cudev;
  PDT_CU_CHECK(d.cuCtxGetDevice(&cudev));
  int vmm_ok = 0, fd_ok = 0;
  PDT_CU_CHECK(d.cuDeviceGetAttribute(&vmm_ok, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cudev));
  PDT_CU_CHECK(d.cuDeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev));
  if (!vmm_ok || !fd_ok) throw std::runtime_error("device lacks VMM / POSIX-fd handle support");

  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = cudev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  PDT_CU_CHECK(d.cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  // multicast binding wants its own (usually larger) granularity: size the heap for both
  size_t mc_gran = 0;
  int mc_ok = 0;
  if (d.multicast_api && world_ > 1) {
    d.cuDeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
    if (mc_ok) {
      CUmulticastObjectProp mp;
      std::memset(&mp, 0, sizeof(mp));
      mp.numDevices = static_cast<unsigned>(world_);
      mp.size = heap_bytes_;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      if (d.cuMulticastGetGranularity(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) mc_gran = 0;
    }
  }
  heap_bytes_ = round_up(heap_bytes_, std::max(gran, std::max<size_t>(mc_gran, 2u << 20)));

  handles_.assign(world_, 0);
  CUmemGenericAllocationHandle mine;
  PDT_CU_CHECK(d.cuMemCreate(&mine, heap_bytes_, &prop, 0));
  handles_[rank_] = mine;
  int my_fd = -1;
  PDT_CU_CHECK(d.cuMemExportToShareableHandle(&my_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::vector<int> fds;
  exchange_fds(store, "heap", my_fd, &fds, timeout);
  ::close(my_fd);

  CUmemAccessDesc access;
  std::memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = cudev;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int r = 0; r < world_; ++r) {
    CUmemGenericAllocationHandle h = mine;
    if (r != rank_) {
      PDT_CU_CHECK(d.cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[r])),
                                                    CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      ::close(fds[r]);
      handles_[r] = h;
    }
    CUdeviceptr va = 0;
    PDT_CU_CHECK(d.cuMemAddressReserve(&va, heap_bytes_, std::max(gran, mc_gran), 0, 0));
    PDT_CU_CHECK(d.cuMemMap(va, heap_bytes_, 0, h, 0));
    PDT_CU_CHECK(d.cuMemSetAccess(va, heap_bytes_, &access, 1));
    peer_base_[r] = reinterpret_cast<char*>(va);
  }
  vmm_ = true;
  if (mc_ok && mc_gran && !(getenv("PDT_SYMM_NO_MULTICAST") && getenv("PDT_SYMM_NO_MULTICAST")[0] == '1')) {
    std::string why;
    bool ok = false;
    try {
      setup_multicast(store, timeout);
      ok = true;
    } catch (const std::exception& e) {
      why = e.what();
    }
    // all-or-nothing across ranks
    store->set("symm/mc_done/" + std::to_string(rank_), ok ? "1" : ("0" + why));
    bool all = true;
    for (int r = 0; r < world_; ++r) all &= (store->get("symm/mc_done/" + std::to_string(r)).substr(0, 1) == "1");
    if (!all) mc_base_ = nullptr;  // keep the mapping alive but never use it
  }
}

void SymmetricHeap::setup_multicast(std::shared_ptr<Store> store, Millis timeout) {
  const DriverApi& d = driver();
  CUdevice cudev;
  PDT_CU_CHECK(d.cuCtxGetDevice(&cudev));
  CUmulticastObjectProp mp;
  std::memset(&mp, 0, sizeof(mp));
  mp.numDevices = static_cast<unsigned>(world_);
  mp.size = heap_bytes_;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mc = 0;
  int root_fd = -1;
  if (rank_ == 0) {
    PDT_CU_CHECK(d.cuMulticastCreate(&mc, &mp));
    PDT_CU_CHECK(d.cuMemExportToShareableHandle(&root_fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  }
  // reuse the pairwise exchange: only rank 0's fd matters, the others pass a dummy (their heap fd
  // is closed already, so hand over stdin's fd number 0 duplicate)
  int send = rank_ == 0 ? root_fd : ::open("/dev/null", O_RDONLY | O_CLOEXEC);
  std::vector<int> fds;
  exchange_fds(store, "mc", send, &fds, timeout);
  ::close(send);
  if (rank_ != 0) {
    PDT_CU_CHECK(d.cuMemImportFromShareableHandle(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[0])),
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  }
  for (int r = 0; r < world_; ++r)
    if (r != rank_ && fds[r] >= 0) ::close(fds[r]);
  PDT_CU_CHECK(d.cuMulticastAddDevice(mc, cudev));
  // every device must be added before anyone binds
  store_barrier(*store, "symm/mc_added", rank_, world_, timeout);
  PDT_CU_CHECK(d.cuMulticastBindMem(mc, 0, handles_[rank_], 0, heap_bytes_, 0));
  CUdeviceptr va = 0;
  size_t mc_gran = 0;
  d.cuMulticastGetGranularity(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
  PDT_CU_CHECK(d.cuMemAddressReserve(&va, heap_bytes_, mc_gran, 0, 0));
  PDT_CU_CHECK(d.cuMemMap(va, heap_bytes_, 0, mc, 0));
  CUmemAccessDesc access;
  std::memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = cudev;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  PDT_CU_CHECK(d.cuMemSetAccess(va, heap_bytes_, &access, 1));
  store_barrier(*store, "symm/mc_bound", rank_, world_, timeout);
  mc_base_ = reinterpret_cast<char*>(va);
  mc_handle_ = mc;
  mc_size_ = heap_bytes_;
}

void SymmetricHeap::setup_ipc(std::shared_ptr<Store> store, Millis timeout) {
  (void)timeout;
  heap_bytes_ = round_up(heap_bytes_, 2u << 20);
  void* mine = nullptr;
  PDT_CUDA_CHECK(cudaMalloc(&mine, heap_bytes_));
  peer_base_[rank_] = static_cast<char*>(mine);
  cudaIpcMemHandle_t h;
  PDT_CUDA_CHECK(cudaIpcGetMemHandle(&h, mine));
  store->set("symm/ipc/" + std::to_string(rank_), std::string(reinterpret_cast<const char*>(&h), sizeof(h)));
  store->set("symm/ipc_dev/" + std::to_string(rank_), std::to_string(device_));
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    std::string blob = store->get("symm/ipc/" + std::to_string(r));
    if (blob.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("symm ipc: bad handle from rank " + std::to_string(r));
    cudaIpcMemHandle_t ph;
    std::memcpy(&ph, blob.data(), sizeof(ph));
    void* p = nullptr;
    PDT_CUDA_CHECK(cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess));
    peer_base_[r] = static_cast<char*>(p);
  }
  vmm_ = false;
  mc_base_ = nullptr;
}

SymmDev SymmetricHeap::dev(int channel) const {
  SymmDev d;
  std::memset(&d, 0, sizeof(d));
  for (int r = 0; r < world_; ++r) d.peer[r] = peer_base_[r];
  d.mc = mc_base_;
  d.flags = nullptr;
  d.flags_off = static_cast<size_t>(channel) * kSymmMaxBlocks * kSymmMaxWorld * sizeof(uint32_t);
  d.epochs = epochs_ + static_cast<size_t>(channel) * kSymmMaxBlocks;
  d.status = status_dev_;
  d.timeout_ns = timeout_ns_;
  d.rank = rank_;
  d.world = world_;
  d.channel = channel;
  return d;
}

size_t SymmetricHeap::staging_off(int channel, int half) const { return staging_off_[channel] + static_cast<size_t>(half) * staging_half_[channel]; }
size_t SymmetricHeap::staging_half_bytes(int channel) const { return staging_half_[channel]; }

void* SymmetricHeap::alloc(size_t nbytes, size_t align) {
  std::lock_guard<std::mutex> g(mu_);
  nbytes = round_up(std::max<size_t>(nbytes, 1), 256);
  align = std::max<size_t>(align, 256);
  for (auto it = free_.begin(); it != free_.end(); ++it) {
    size_t start = round_up(it->first, align);
    size_t pad = start - it->first;
    if (it->second < pad + nbytes) continue;
    size_t blk_off = it->first, blk_sz = it->second;
    free_.erase(it);
    if (pad) free_[blk_off] = pad;
    size_t tail = blk_sz - pad - nbytes;
    if (tail) free_[start + nbytes] = tail;
    used_[start] = nbytes;
    return peer_base_[rank_] + start;
  }
  throw std::runtime_error("symmetric heap exhausted: requested " + std::to_string(nbytes) + " bytes; raise PDT_SYMM_HEAP_MB");
}

void SymmetricHeap::free(void* p) {
  // Deleters run whenever Python drops the last reference — a moment that differs between ranks (refcounts, cyclic GC).
  // The block is only *parked* here; it returns to the allocator at the next collective point (SymmComm::alloc_flat),
  // where the ranks agree on the set of blocks everybody has released, so the first-fit state stays identical on all
  // ranks and never recycles memory a comm-stream kernel may still be using.
  std::lock_guard<std::mutex> g(mu_);
  size_t off = static_cast<char*>(p) - peer_base_[rank_];
  if (used_.find(off) == used_.end()) return;
  if (world_ == 1) {
    release_locked(off);
    return;
  }
  pending_.insert(off);
}

std::vector<size_t> SymmetricHeap::pending_frees() const {
  std::lock_guard<std::mutex> g(mu_);
  return std::vector<size_t>(pending_.begin(), pending_.end());
}

void SymmetricHeap::apply_frees(const std::vector<size_t>& offsets) {
  std::lock_guard<std::mutex> g(mu_);
  for (size_t off : offsets) {
    if (pending_.erase(off)) release_locked(off);
  }
}

void SymmetricHeap::release_locked(size_t off) {
  auto it = used_.find(off);
  if (it == used_.end()) return;
  size_t sz = it->second;
  used_.erase(it);
  auto ins = free_.emplace(off, sz).first;
  // coalesce with the right neighbour, then the left one
  auto right = std::next(ins);
  if (right != free_.end() && ins->first + ins->second == right->first) {
    ins->second += right->second;
    free_.erase(right);
  }
  if (ins != free_.begin()) {
    auto left = std::prev(ins);
    if (left->first + left->second == ins->first) {
      left->second += ins->second;
      free_.erase(ins);
    }
  }
}

bool SymmetricHeap::contains(const void* p, size_t nbytes) const {
  const char* c = static_cast<const char*>(p);
  return c >= peer_base_[rank_] + user_off_ && c + nbytes <= peer_base_[rank_] + heap_bytes_;
}

size_t SymmetricHeap::user_bytes_in_use() const {
  std::lock_guard<std::mutex> g(mu_);
  size_t n = 0;
  for (auto& kv : used_) n += kv.second;
  return n;
}

}  // namespace pdt
