#include "store.h"

#include <fcntl.h>
#include <poll.h>
#include <sys/file.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>

namespace pdt {

// ---------------------------------------------------------------------------------------
// KVState
int64_t KVState::add(const std::string& k, int64_t d) {
  int64_t cur = 0;
  auto it = kv_.find(k);
  if (it != kv_.end() && !it->second.empty()) cur = std::stoll(it->second);
  cur += d;
  kv_[k] = std::to_string(cur);
  return cur;
}

Bytes KVState::compare_set(const std::string& k, const Bytes& expected, const Bytes& desired) {
  auto it = kv_.find(k);
  if (it == kv_.end()) {
    if (expected.empty()) {
      kv_[k] = desired;
      return desired;
    }
    return expected;  // c10d semantics: missing key + non-empty expectation → echo expectation
  }
  if (it->second == expected) it->second = desired;
  return it->second;
}

bool KVState::qpop(const std::string& k, Bytes* out) {
  auto it = q_.find(k);
  if (it == q_.end() || it->second.empty()) return false;
  *out = std::move(it->second.front());
  it->second.pop_front();
  return true;
}

int64_t KVState::qlen(const std::string& k) const {
  auto it = q_.find(k);
  return it == q_.end() ? 0 : static_cast<int64_t>(it->second.size());
}

// ---------------------------------------------------------------------------------------
// HashStore
void HashStore::set(const std::string& key, const Bytes& value) {
  { std::lock_guard<std::mutex> g(mu_); st_.set(key, value); }
  cv_.notify_all();
}
Bytes HashStore::get(const std::string& key) {
  std::unique_lock<std::mutex> g(mu_);
  if (!cv_.wait_for(g, timeout_, [&] { return st_.has(key); }))
    throw TimeoutError("HashStore.get('" + key + "') timed out");
  return st_.at(key);
}
int64_t HashStore::add(const std::string& key, int64_t delta) {
  int64_t v;
  { std::lock_guard<std::mutex> g(mu_); v = st_.add(key, delta); }
  cv_.notify_all();
  return v;
}
Bytes HashStore::compare_set(const std::string& key, const Bytes& e, const Bytes& d) {
  Bytes v;
  { std::lock_guard<std::mutex> g(mu_); v = st_.compare_set(key, e, d); }
  cv_.notify_all();
  return v;
}
void HashStore::wait(const std::vector<std::string>& keys, Millis timeout) {
  std::unique_lock<std::mutex> g(mu_);
  bool ok = cv_.wait_for(g, timeout, [&] {
    for (auto& k : keys) if (!st_.has(k)) return false;
    return true;
  });
  if (!ok) throw TimeoutError("HashStore.wait timed out");
}
bool HashStore::check(const std::vector<std::string>& keys) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto& k : keys) if (!st_.has(k)) return false;
  return true;
}
bool HashStore::delete_key(const std::string& key) { std::lock_guard<std::mutex> g(mu_); return st_.erase(key); }
int64_t HashStore::num_keys() { std::lock_guard<std::mutex> g(mu_); return st_.size(); }
void HashStore::append(const std::string& key, const Bytes& value) {
  { std::lock_guard<std::mutex> g(mu_); st_.append(key, value); }
  cv_.notify_all();
}
std::vector<Bytes> HashStore::multi_get(const std::vector<std::string>& keys) {
  wait(keys, timeout_);
  std::lock_guard<std::mutex> g(mu_);
  std::vector<Bytes> out;
  for (auto& k : keys) out.push_back(st_.at(k));
  return out;
}
void HashStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  if (keys.size() != values.size()) throw std::invalid_argument("multi_set: keys/values length mismatch");
  { std::lock_guard<std::mutex> g(mu_); for (size_t i = 0; i < keys.size(); ++i) st_.set(keys[i], values[i]); }
  cv_.notify_all();
}
void HashStore::queue_push(const std::string& key, const Bytes& value) {
  { std::lock_guard<std::mutex> g(mu_); st_.qpush(key, value); }
  cv_.notify_all();
}
Bytes HashStore::queue_pop(const std::string& key, bool block) {
  std::unique_lock<std::mutex> g(mu_);
  Bytes out;
  if (st_.qpop(key, &out)) return out;
  if (!block) throw std::out_of_range("queue '" + key + "' is empty");
  bool ok = cv_.wait_for(g, timeout_, [&] { return st_.qpop(key, &out); });
  if (!ok) throw TimeoutError("HashStore.queue_pop('" + key + "') timed out");
  return out;
}
int64_t HashStore::queue_len(const std::string& key) { std::lock_guard<std::mutex> g(mu_); return st_.qlen(key); }

// ---------------------------------------------------------------------------------------
// FileStore: log records  [u8 op][u32 klen][u32 vlen][k][v]
namespace {
enum FileOp : uint8_t { F_SET = 1, F_DEL = 2, F_APPEND = 3, F_QPUSH = 4, F_QPOP = 5 };
}

struct FileStore::Locked {
  FileStore& s;
  int fd;
  explicit Locked(FileStore& st) : s(st) {
    fd = ::open(s.path_.c_str(), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
    if (fd < 0) throw std::runtime_error(errno_str("FileStore open " + s.path_));
    while (::flock(fd, LOCK_EX) != 0) {
      if (errno != EINTR) { ::close(fd); throw std::runtime_error(errno_str("flock")); }
    }
    s.replay_locked(fd);
  }
  ~Locked() { ::flock(fd, LOCK_UN); ::close(fd); }
};

FileStore::FileStore(std::string path, int world_size) : path_(std::move(path)), world_size_(world_size) {
  Locked l(*this);
  (void)l;
}
FileStore::~FileStore() = default;

void FileStore::replay_locked(int fd) {
  struct stat sb;
  if (::fstat(fd, &sb) != 0) throw std::runtime_error(errno_str("fstat"));
  while (pos_ < sb.st_size) {
    uint8_t hdr[9];
    if (::pread(fd, hdr, 9, pos_) != 9) break;
    uint32_t kl, vl;
    std::memcpy(&kl, hdr + 1, 4);
    std::memcpy(&vl, hdr + 5, 4);
    if (pos_ + 9 + static_cast<off_t>(kl) + static_cast<off_t>(vl) > sb.st_size) break;  // torn tail
    std::string k(kl, '\0'), v(vl, '\0');
    if (kl && ::pread(fd, &k[0], kl, pos_ + 9) != static_cast<ssize_t>(kl)) break;
    if (vl && ::pread(fd, &v[0], vl, pos_ + 9 + kl) != static_cast<ssize_t>(vl)) break;
    switch (hdr[0]) {
      case F_SET: st_.set(k, v); break;
      case F_DEL: st_.erase(k); break;
      case F_APPEND: st_.append(k, v); break;
      case F_QPUSH: st_.qpush(k, v); break;
      case F_QPOP: { Bytes tmp; st_.qpop(k, &tmp); break; }
      default: break;
    }
    pos_ += 9 + kl + vl;
  }
}

void FileStore::log_locked(int fd, uint8_t op, const std::string& k, const Bytes& v) {
  std::string rec(9, '\0');
  rec[0] = static_cast<char>(op);
  uint32_t kl = static_cast<uint32_t>(k.size()), vl = static_cast<uint32_t>(v.size());
  std::memcpy(&rec[1], &kl, 4);
  std::memcpy(&rec[5], &vl, 4);
  rec += k;
  rec += v;
  if (::pwrite(fd, rec.data(), rec.size(), pos_) != static_cast<ssize_t>(rec.size()))
    throw std::runtime_error(errno_str("FileStore write"));
  pos_ += static_cast<off_t>(rec.size());
}

void FileStore::set(const std::string& key, const Bytes& value) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  st_.set(key, value);
  log_locked(l.fd, F_SET, key, value);
}
Bytes FileStore::get(const std::string& key) {
  wait({key}, timeout_);
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  return st_.at(key);
}
int64_t FileStore::add(const std::string& key, int64_t delta) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  int64_t v = st_.add(key, delta);
  log_locked(l.fd, F_SET, key, std::to_string(v));
  return v;
}
Bytes FileStore::compare_set(const std::string& key, const Bytes& e, const Bytes& d) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  bool had = st_.has(key);
  Bytes before = had ? st_.at(key) : Bytes();
  Bytes v = st_.compare_set(key, e, d);
  if (st_.has(key) && (!had || st_.at(key) != before)) log_locked(l.fd, F_SET, key, st_.at(key));
  return v;
}
void FileStore::wait(const std::vector<std::string>& keys, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  int sleep_ms = 1;
  while (true) {
    if (check(keys)) return;
    if (Clock::now() >= deadline) throw TimeoutError("FileStore.wait timed out");
    ::usleep(sleep_ms * 1000);
    if (sleep_ms < 32) sleep_ms *= 2;
  }
}
bool FileStore::check(const std::vector<std::string>& keys) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  for (auto& k : keys) if (!st_.has(k)) return false;
  return true;
}
bool FileStore::delete_key(const std::string& key) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  bool r = st_.erase(key);
  if (r) log_locked(l.fd, F_DEL, key, "");
  return r;
}
int64_t FileStore::num_keys() {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  return st_.size();
}
void FileStore::append(const std::string& key, const Bytes& value) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  st_.append(key, value);
  log_locked(l.fd, F_APPEND, key, value);
}
std::vector<Bytes> FileStore::multi_get(const std::vector<std::string>& keys) {
  wait(keys, timeout_);
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  std::vector<Bytes> out;
  for (auto& k : keys) out.push_back(st_.at(k));
  return out;
}
void FileStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  if (keys.size() != values.size()) throw std::invalid_argument("multi_set: keys/values length mismatch");
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  for (size_t i = 0; i < keys.size(); ++i) {
    st_.set(keys[i], values[i]);
    log_locked(l.fd, F_SET, keys[i], values[i]);
  }
}
void FileStore::queue_push(const std::string& key, const Bytes& value) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  st_.qpush(key, value);
  log_locked(l.fd, F_QPUSH, key, value);
}
Bytes FileStore::queue_pop(const std::string& key, bool block) {
  auto deadline = Clock::now() + timeout_;
  while (true) {
    {
      std::lock_guard<std::mutex> g(mu_);
      Locked l(*this);
      Bytes out;
      if (st_.qpop(key, &out)) {
        log_locked(l.fd, F_QPOP, key, "");
        return out;
      }
    }
    if (!block) throw std::out_of_range("queue '" + key + "' is empty");
    if (Clock::now() >= deadline) throw TimeoutError("FileStore.queue_pop timed out");
    ::usleep(2000);
  }
}
int64_t FileStore::queue_len(const std::string& key) {
  std::lock_guard<std::mutex> g(mu_);
  Locked l(*this);
  return st_.qlen(key);
}

// ---------------------------------------------------------------------------------------
// TCP wire format.
//   request : u8 op | u32 timeout_ms | u32 nargs | nargs × (u32 len | bytes)
//   response: u8 status | u32 nargs | nargs × (u32 len | bytes)
namespace {
enum Op : uint8_t {
  OP_SET = 1, OP_GET, OP_ADD, OP_CAS, OP_WAIT, OP_CHECK, OP_DEL, OP_NUMKEYS, OP_APPEND,
  OP_MGET, OP_MSET, OP_QPUSH, OP_QPOP, OP_QLEN, OP_PING
};
enum Status : uint8_t { ST_OK = 0, ST_TIMEOUT = 1, ST_EMPTY = 2, ST_BAD = 3 };

void put_u32(std::string& s, uint32_t v) { s.append(reinterpret_cast<const char*>(&v), 4); }
void put_arg(std::string& s, const Bytes& b) { put_u32(s, static_cast<uint32_t>(b.size())); s += b; }

std::string encode_response(uint8_t status, const std::vector<Bytes>& args) {
  std::string out;
  out.push_back(static_cast<char>(status));
  put_u32(out, static_cast<uint32_t>(args.size()));
  for (auto& a : args) put_arg(out, a);
  return out;
}
}  // namespace

struct TCPStoreServer::Conn {
  Fd fd;
  std::string in;
  bool blocked = false;  // a waiter is registered: do not parse further frames until answered
};

struct TCPStoreServer::Waiter {
  int fd;
  uint8_t op;
  std::vector<std::string> keys;
  Clock::time_point deadline;
};

TCPStoreServer::TCPStoreServer(const std::string& host, int port) {
  listen_fd_ = tcp_listen(host, port, &port_);
  set_nonblocking(listen_fd_.get(), true);  // the accept loop drains until EAGAIN
  int p[2];
  if (::pipe2(p, O_CLOEXEC | O_NONBLOCK) != 0) throw std::runtime_error(errno_str("pipe2"));
  wake_r_ = p[0];
  wake_w_ = p[1];
  thread_ = std::thread([this] { loop(); });
}

TCPStoreServer::~TCPStoreServer() { shutdown(); }

void TCPStoreServer::shutdown() {
  bool was = stop_.exchange(true);
  if (!was) {
    char c = 'x';
    (void)!::write(wake_w_, &c, 1);
  }
  if (thread_.joinable()) thread_.join();
  if (wake_r_ >= 0) { ::close(wake_r_); wake_r_ = -1; }
  if (wake_w_ >= 0) { ::close(wake_w_); wake_w_ = -1; }
  conns_.clear();
  listen_fd_.reset();
}

static bool send_response(int fd, const std::string& resp) {
  try {
    send_all(fd, resp.data(), resp.size(), Millis(30000));
    return true;
  } catch (const std::exception&) {
    return false;
  }
}

void TCPStoreServer::serve_waiters() {
  auto now = Clock::now();
  for (size_t i = 0; i < waiters_.size();) {
    Waiter& w = waiters_[i];
    auto it = conns_.find(w.fd);
    if (it == conns_.end()) { waiters_.erase(waiters_.begin() + i); continue; }
    bool done = false;
    std::string resp;
    if (w.op == OP_QPOP) {
      Bytes v;
      if (st_.qpop(w.keys[0], &v)) { resp = encode_response(ST_OK, {v}); done = true; }
    } else {
      bool all = true;
      for (auto& k : w.keys) if (!st_.has(k)) { all = false; break; }
      if (all) {
        std::vector<Bytes> vals;
        if (w.op == OP_GET || w.op == OP_MGET) for (auto& k : w.keys) vals.push_back(st_.at(k));
        resp = encode_response(ST_OK, vals);
        done = true;
      }
    }
    if (!done && now >= w.deadline) { resp = encode_response(ST_TIMEOUT, {}); done = true; }
    if (done) {
      int fd = w.fd;
      waiters_.erase(waiters_.begin() + i);
      it->second->blocked = false;
      if (!send_response(fd, resp)) { conns_.erase(fd); continue; }
      // frames that queued up behind the blocked request
      while (!it->second->blocked && handle_frame(*it->second)) {}
      continue;
    }
    ++i;
  }
}

// Parses and executes at most one frame from c.in. Returns true if a frame was consumed.
bool TCPStoreServer::handle_frame(Conn& c) {
  const std::string& in = c.in;
  if (in.size() < 9) return false;
  uint8_t op = static_cast<uint8_t>(in[0]);
  uint32_t timeout_ms, nargs;
  std::memcpy(&timeout_ms, &in[1], 4);
  std::memcpy(&nargs, &in[5], 4);
  size_t off = 9;
  std::vector<Bytes> args;
  args.reserve(nargs);
  for (uint32_t i = 0; i < nargs; ++i) {
    if (in.size() < off + 4) return false;
    uint32_t len;
    std::memcpy(&len, &in[off], 4);
    off += 4;
    if (in.size() < off + len) return false;
    args.emplace_back(in.data() + off, len);
    off += len;
  }
  c.in.erase(0, off);
  int fd = c.fd.get();
  auto reply = [&](uint8_t st, const std::vector<Bytes>& a) { send_response(fd, encode_response(st, a)); };
  auto need = [&](size_t n) { if (args.size() < n) { reply(ST_BAD, {}); return false; } return true; };
  bool wake = false;
  switch (op) {
    case OP_SET: if (!need(2)) break; st_.set(args[0], args[1]); reply(ST_OK, {}); wake = true; break;
    case OP_ADD: {
      if (!need(2)) break;
      int64_t v = st_.add(args[0], std::stoll(args[1]));
      reply(ST_OK, {std::to_string(v)});
      wake = true;
      break;
    }
    case OP_CAS: if (!need(3)) break; reply(ST_OK, {st_.compare_set(args[0], args[1], args[2])}); wake = true; break;
    case OP_CHECK: {
      bool all = true;
      for (auto& k : args) if (!st_.has(k)) { all = false; break; }
      reply(ST_OK, {all ? "1" : "0"});
      break;
    }
    case OP_DEL: if (!need(1)) break; reply(ST_OK, {st_.erase(args[0]) ? "1" : "0"}); break;
    case OP_NUMKEYS: reply(ST_OK, {std::to_string(st_.size())}); break;
    case OP_APPEND: if (!need(2)) break; st_.append(args[0], args[1]); reply(ST_OK, {}); wake = true; break;
    case OP_MSET: {
      if (args.size() % 2) { reply(ST_BAD, {}); break; }
      for (size_t i = 0; i < args.size(); i += 2) st_.set(args[i], args[i + 1]);
      reply(ST_OK, {});
      wake = true;
      break;
    }
    case OP_QPUSH: if (!need(2)) break; st_.qpush(args[0], args[1]); reply(ST_OK, {}); wake = true; break;
    case OP_QLEN: if (!need(1)) break; reply(ST_OK, {std::to_string(st_.qlen(args[0]))}); break;
    case OP_PING: reply(ST_OK, args); break;
    case OP_QPOP: {
      if (!need(2)) break;
      Bytes v;
      if (st_.qpop(args[0], &v)) { reply(ST_OK, {v}); break; }
      if (args[1] == "0") { reply(ST_EMPTY, {}); break; }
      c.blocked = true;
      waiters_.push_back({fd, op, {args[0]}, Clock::now() + Millis(timeout_ms)});
      break;
    }
    case OP_GET: case OP_MGET: case OP_WAIT: {
      c.blocked = true;
      waiters_.push_back({fd, op, args, Clock::now() + Millis(timeout_ms)});
      wake = true;  // may be immediately satisfiable
      break;
    }
    default: reply(ST_BAD, {}); break;
  }
  if (wake) dirty_ = true;  // the main loop re-runs serve_waiters(); never recurse from here
  return true;
}

void TCPStoreServer::loop() {
  std::vector<struct pollfd> pfds;
  while (!stop_.load()) {
    pfds.clear();
    pfds.push_back({listen_fd_.get(), POLLIN, 0});
    pfds.push_back({wake_r_, POLLIN, 0});
    for (auto& kv : conns_) pfds.push_back({kv.first, POLLIN, 0});
    int timeout_ms = 1000;
    auto now = Clock::now();
    for (auto& w : waiters_) {
      auto left = std::chrono::duration_cast<Millis>(w.deadline - now).count();
      timeout_ms = static_cast<int>(std::max<int64_t>(0, std::min<int64_t>(timeout_ms, left)));
    }
    int rc = ::poll(pfds.data(), pfds.size(), timeout_ms);
    if (rc < 0) {
      if (errno == EINTR) continue;
      break;
    }
    if (stop_.load()) break;
    if (pfds[0].revents & POLLIN) {
      while (true) {
        int fd = ::accept4(listen_fd_.get(), nullptr, nullptr, SOCK_CLOEXEC | SOCK_NONBLOCK);
        if (fd < 0) break;
        set_nodelay(fd);
        auto c = std::make_unique<Conn>();
        c->fd.reset(fd);
        conns_[fd] = std::move(c);
      }
    }
    if (pfds[1].revents & POLLIN) {
      char buf[64];
      while (::read(wake_r_, buf, sizeof(buf)) > 0) {}
    }
    for (size_t i = 2; i < pfds.size(); ++i) {
      if (!(pfds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
      auto it = conns_.find(pfds[i].fd);
      if (it == conns_.end()) continue;
      Conn& c = *it->second;
      bool closed = false;
      char buf[65536];
      while (true) {
        ssize_t r = ::recv(c.fd.get(), buf, sizeof(buf), MSG_DONTWAIT);
        if (r > 0) { c.in.append(buf, static_cast<size_t>(r)); continue; }
        if (r == 0) { closed = true; break; }
        if (errno == EINTR) continue;
        if (errno == EAGAIN || errno == EWOULDBLOCK) break;
        closed = true;
        break;
      }
      while (!c.blocked && handle_frame(c)) {}
      if (closed) {
        int fd = pfds[i].fd;
        waiters_.erase(std::remove_if(waiters_.begin(), waiters_.end(), [fd](const Waiter& w) { return w.fd == fd; }),
                       waiters_.end());
        conns_.erase(fd);
      }
    }
    do {
      dirty_ = false;
      serve_waiters();
    } while (dirty_);
  }
}

// ---------------------------------------------------------------------------------------
// TCPStore client
TCPStore::TCPStore(const std::string& host, int port, int world_size, bool is_master, Millis timeout,
                   bool wait_for_workers)
    : host_(host), port_(port) {
  timeout_ = timeout;
  if (is_master) {
    server_ = std::make_unique<TCPStoreServer>(host == "localhost" ? "" : host, port);
    port_ = server_->port();
  }
  idle_.push_back(tcp_connect(host_, port_, timeout_));
  if (world_size > 0) {
    const std::string kInit = "__pdt_store_init__/workers";
    add(kInit, 1);
    if (is_master && wait_for_workers) {
      auto deadline = Clock::now() + timeout_;
      while (add(kInit, 0) < world_size) {
        if (Clock::now() >= deadline)
          throw TimeoutError("TCPStore master: timed out waiting for " + std::to_string(world_size) + " workers");
        ::usleep(1000);
      }
    }
  }
}

TCPStore::~TCPStore() {
  idle_.clear();
  server_.reset();
}

std::vector<Bytes> TCPStore::call(uint8_t op, const std::vector<Bytes>& args, Millis timeout, uint8_t* status) {
  // One request in flight per connection. A blocking get()/wait() must not stall set() from
  // another thread of the same process, so connections are pooled and dialled on demand.
  Fd conn;
  {
    std::lock_guard<std::mutex> g(mu_);
    if (!idle_.empty()) {
      conn = std::move(idle_.back());
      idle_.pop_back();
    }
  }
  if (!conn.valid()) conn = tcp_connect(host_, port_, timeout_);
  std::string req;
  req.push_back(static_cast<char>(op));
  put_u32(req, static_cast<uint32_t>(std::min<int64_t>(timeout.count(), 0x7fffffff)));
  put_u32(req, static_cast<uint32_t>(args.size()));
  for (auto& a : args) put_arg(req, a);
  // The server enforces `timeout` for blocking ops; the socket deadline is a safety net on top.
  Millis io = timeout + Millis(15000);
  send_all(conn.get(), req.data(), req.size(), io);
  uint8_t hdr[5];
  recv_all(conn.get(), hdr, 5, io);
  uint32_t n;
  std::memcpy(&n, hdr + 1, 4);
  std::vector<Bytes> out;
  out.reserve(n);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t len;
    recv_all(conn.get(), &len, 4, io);
    Bytes b(len, '\0');
    if (len) recv_all(conn.get(), &b[0], len, io);
    out.push_back(std::move(b));
  }
  {
    // frame fully consumed: the connection is clean and can be reused (an exception above
    // drops it instead)
    std::lock_guard<std::mutex> g(mu_);
    if (idle_.size() < 8) idle_.push_back(std::move(conn));
  }
  if (status) *status = hdr[0];
  else if (hdr[0] == ST_TIMEOUT) throw TimeoutError("TCPStore: operation timed out after " + std::to_string(timeout.count()) + " ms");
  else if (hdr[0] != ST_OK) throw std::runtime_error("TCPStore: server rejected request (status " + std::to_string(hdr[0]) + ")");
  return out;
}

void TCPStore::set(const std::string& key, const Bytes& value) { call(OP_SET, {key, value}, timeout_); }
Bytes TCPStore::get(const std::string& key) {
  uint8_t st;
  auto r = call(OP_GET, {key}, timeout_, &st);
  if (st == ST_TIMEOUT) throw TimeoutError("TCPStore.get('" + key + "') timed out after " + std::to_string(timeout_.count()) + " ms");
  if (st != ST_OK || r.empty()) throw std::runtime_error("TCPStore.get failed");
  return r[0];
}
int64_t TCPStore::add(const std::string& key, int64_t delta) {
  return std::stoll(call(OP_ADD, {key, std::to_string(delta)}, timeout_).at(0));
}
Bytes TCPStore::compare_set(const std::string& key, const Bytes& e, const Bytes& d) {
  return call(OP_CAS, {key, e, d}, timeout_).at(0);
}
void TCPStore::wait(const std::vector<std::string>& keys, Millis timeout) {
  uint8_t st;
  call(OP_WAIT, keys, timeout, &st);
  if (st == ST_TIMEOUT) {
    std::string ks;
    for (auto& k : keys) ks += (ks.empty() ? "" : ", ") + k;
    throw TimeoutError("TCPStore.wait([" + ks + "]) timed out after " + std::to_string(timeout.count()) + " ms");
  }
  if (st != ST_OK) throw std::runtime_error("TCPStore.wait failed");
}
bool TCPStore::check(const std::vector<std::string>& keys) { return call(OP_CHECK, keys, timeout_).at(0) == "1"; }
bool TCPStore::delete_key(const std::string& key) { return call(OP_DEL, {key}, timeout_).at(0) == "1"; }
int64_t TCPStore::num_keys() { return std::stoll(call(OP_NUMKEYS, {}, timeout_).at(0)); }
void TCPStore::append(const std::string& key, const Bytes& value) { call(OP_APPEND, {key, value}, timeout_); }
std::vector<Bytes> TCPStore::multi_get(const std::vector<std::string>& keys) {
  uint8_t st;
  auto r = call(OP_MGET, keys, timeout_, &st);
  if (st == ST_TIMEOUT) throw TimeoutError("TCPStore.multi_get timed out");
  if (st != ST_OK) throw std::runtime_error("TCPStore.multi_get failed");
  return r;
}
void TCPStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  if (keys.size() != values.size()) throw std::invalid_argument("multi_set: keys/values length mismatch");
  std::vector<Bytes> args;
  for (size_t i = 0; i < keys.size(); ++i) { args.push_back(keys[i]); args.push_back(values[i]); }
  call(OP_MSET, args, timeout_);
}
void TCPStore::queue_push(const std::string& key, const Bytes& value) { call(OP_QPUSH, {key, value}, timeout_); }
Bytes TCPStore::queue_pop(const std::string& key, bool block) {
  uint8_t st;
  auto r = call(OP_QPOP, {key, block ? "1" : "0"}, timeout_, &st);
  if (st == ST_EMPTY) throw std::out_of_range("queue '" + key + "' is empty");
  if (st == ST_TIMEOUT) throw TimeoutError("TCPStore.queue_pop('" + key + "') timed out");
  if (st != ST_OK || r.empty()) throw std::runtime_error("TCPStore.queue_pop failed");
  return r[0];
}
int64_t TCPStore::queue_len(const std::string& key) { return std::stoll(call(OP_QLEN, {key}, timeout_).at(0)); }
void TCPStore::ping() { call(OP_PING, {"ping"}, timeout_); }

// ---------------------------------------------------------------------------------------
void store_barrier(Store& store, const std::string& name, int rank, int world_size, Millis timeout) {
  // Generation counter makes the same barrier name reusable: arrival i belongs to
  // generation floor((i-1)/world_size).
  (void)rank;
  int64_t arrival = store.add("__barrier__/" + name + "/count", 1);
  int64_t gen = (arrival - 1) / world_size;
  std::string done_key = "__barrier__/" + name + "/done/" + std::to_string(gen);
  if (arrival == (gen + 1) * world_size) store.set(done_key, "1");
  store.wait({done_key}, timeout);
}

}  // namespace pdt
