// Key/value rendezvous stores.
//
// Behavioural parity target: the c10d Store family the reference reaches through
// dist.init_process_group(init_method='tcp://...') (ref: ddp_example.py:50,110;
// torch/include/torch/csrc/distributed/c10d/{Store,TCPStore,PrefixStore,HashStore,FileStore}.hpp).
// The design is our own: one poll()-driven server thread, length-prefixed binary frames,
// server-side wait queues with server-enforced deadlines (a timed-out waiter can never
// desynchronise the connection), and a monotonically increasing "generation" for barriers.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../common/net.h"

namespace pdt {

using Bytes = std::string;  // arbitrary binary payload

class Store {
 public:
  virtual ~Store() = default;
  virtual void set(const std::string& key, const Bytes& value) = 0;
  // Blocks until the key exists (or timeout → TimeoutError).
  virtual Bytes get(const std::string& key) = 0;
  // Atomic add on an int64 counter stored as decimal text; missing key counts as 0.
  virtual int64_t add(const std::string& key, int64_t delta) = 0;
  // If key is missing and expected is empty → set desired; if current==expected → set desired.
  // Returns the value now stored.
  virtual Bytes compare_set(const std::string& key, const Bytes& expected, const Bytes& desired) = 0;
  virtual void wait(const std::vector<std::string>& keys) { wait(keys, timeout_); }
  virtual void wait(const std::vector<std::string>& keys, Millis timeout) = 0;
  virtual bool check(const std::vector<std::string>& keys) = 0;
  virtual bool delete_key(const std::string& key) = 0;
  virtual int64_t num_keys() = 0;
  virtual void append(const std::string& key, const Bytes& value) = 0;
  virtual std::vector<Bytes> multi_get(const std::vector<std::string>& keys) = 0;
  virtual void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) = 0;
  virtual void queue_push(const std::string& key, const Bytes& value) = 0;
  virtual Bytes queue_pop(const std::string& key, bool block) = 0;
  virtual int64_t queue_len(const std::string& key) = 0;

  Millis timeout() const { return timeout_; }
  virtual void set_timeout(Millis t) { timeout_ = t; }

 protected:
  Millis timeout_{300 * 1000};
};

// ---------------------------------------------------------------------------------------
// In-memory state machine shared by HashStore (in-process) and the TCP server.
// Not thread safe by itself.
class KVState {
 public:
  void set(const std::string& k, const Bytes& v) { kv_[k] = v; }
  bool has(const std::string& k) const { return kv_.count(k) != 0; }
  const Bytes& at(const std::string& k) const { return kv_.at(k); }
  int64_t add(const std::string& k, int64_t d);
  Bytes compare_set(const std::string& k, const Bytes& expected, const Bytes& desired);
  bool erase(const std::string& k) { return kv_.erase(k) != 0; }
  int64_t size() const { return static_cast<int64_t>(kv_.size()); }
  void append(const std::string& k, const Bytes& v) { kv_[k] += v; }
  void qpush(const std::string& k, const Bytes& v) { q_[k].push_back(v); }
  bool qpop(const std::string& k, Bytes* out);
  int64_t qlen(const std::string& k) const;

 private:
  std::unordered_map<std::string, Bytes> kv_;
  std::unordered_map<std::string, std::deque<Bytes>> q_;
};

// ---------------------------------------------------------------------------------------
class HashStore : public Store {
 public:
  void set(const std::string& key, const Bytes& value) override;
  Bytes get(const std::string& key) override;
  int64_t add(const std::string& key, int64_t delta) override;
  Bytes compare_set(const std::string& key, const Bytes& expected, const Bytes& desired) override;
  using Store::wait;
  void wait(const std::vector<std::string>& keys, Millis timeout) override;
  bool check(const std::vector<std::string>& keys) override;
  bool delete_key(const std::string& key) override;
  int64_t num_keys() override;
  void append(const std::string& key, const Bytes& value) override;
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void queue_push(const std::string& key, const Bytes& value) override;
  Bytes queue_pop(const std::string& key, bool block) override;
  int64_t queue_len(const std::string& key) override;

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  KVState st_;
};

// ---------------------------------------------------------------------------------------
// File-backed store (init_method='file://...'): an append-only log of (op,key,value)
// records guarded by flock(); every operation replays the unseen tail.
class FileStore : public Store {
 public:
  FileStore(std::string path, int world_size);
  ~FileStore() override;
  void set(const std::string& key, const Bytes& value) override;
  Bytes get(const std::string& key) override;
  int64_t add(const std::string& key, int64_t delta) override;
  Bytes compare_set(const std::string& key, const Bytes& expected, const Bytes& desired) override;
  using Store::wait;
  void wait(const std::vector<std::string>& keys, Millis timeout) override;
  bool check(const std::vector<std::string>& keys) override;
  bool delete_key(const std::string& key) override;
  int64_t num_keys() override;
  void append(const std::string& key, const Bytes& value) override;
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void queue_push(const std::string& key, const Bytes& value) override;
  Bytes queue_pop(const std::string& key, bool block) override;
  int64_t queue_len(const std::string& key) override;
  const std::string& path() const { return path_; }

 private:
  struct Locked;  // RAII flock + replay
  void replay_locked(int fd);
  void log_locked(int fd, uint8_t op, const std::string& k, const Bytes& v);
  std::string path_;
  int world_size_;
  std::mutex mu_;
  KVState st_;
  off_t pos_ = 0;
};

// ---------------------------------------------------------------------------------------
class PrefixStore : public Store {
 public:
  PrefixStore(std::string prefix, std::shared_ptr<Store> base)
      : prefix_(std::move(prefix)), base_(std::move(base)) { timeout_ = base_->timeout(); }
  void set(const std::string& key, const Bytes& value) override { base_->set(k(key), value); }
  Bytes get(const std::string& key) override { return base_->get(k(key)); }
  int64_t add(const std::string& key, int64_t delta) override { return base_->add(k(key), delta); }
  Bytes compare_set(const std::string& key, const Bytes& e, const Bytes& d) override {
    return base_->compare_set(k(key), e, d);
  }
  using Store::wait;
  void wait(const std::vector<std::string>& keys, Millis timeout) override { base_->wait(ks(keys), timeout); }
  bool check(const std::vector<std::string>& keys) override { return base_->check(ks(keys)); }
  bool delete_key(const std::string& key) override { return base_->delete_key(k(key)); }
  int64_t num_keys() override { return base_->num_keys(); }
  void append(const std::string& key, const Bytes& value) override { base_->append(k(key), value); }
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override { return base_->multi_get(ks(keys)); }
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override {
    base_->multi_set(ks(keys), values);
  }
  void queue_push(const std::string& key, const Bytes& value) override { base_->queue_push(k(key), value); }
  Bytes queue_pop(const std::string& key, bool block) override { return base_->queue_pop(k(key), block); }
  int64_t queue_len(const std::string& key) override { return base_->queue_len(k(key)); }
  void set_timeout(Millis t) override { timeout_ = t; base_->set_timeout(t); }
  const std::string& prefix() const { return prefix_; }
  std::shared_ptr<Store> underlying() const { return base_; }

 private:
  std::string k(const std::string& key) const { return prefix_ + "/" + key; }
  std::vector<std::string> ks(const std::vector<std::string>& keys) const {
    std::vector<std::string> out;
    out.reserve(keys.size());
    for (auto& x : keys) out.push_back(k(x));
    return out;
  }
  std::string prefix_;
  std::shared_ptr<Store> base_;
};

// ---------------------------------------------------------------------------------------
class TCPStoreServer {
 public:
  TCPStoreServer(const std::string& host, int port);
  ~TCPStoreServer();
  int port() const { return port_; }
  void shutdown();

 private:
  struct Conn;
  struct Waiter;
  void loop();
  bool handle_frame(Conn& c);            // returns false to drop the connection
  void serve_waiters();
  Fd listen_fd_;
  int wake_r_ = -1, wake_w_ = -1;        // self-pipe
  int port_ = 0;
  std::atomic<bool> stop_{false};
  bool dirty_ = false;                   // state changed since the last waiter scan
  std::thread thread_;
  KVState st_;
  std::map<int, std::unique_ptr<Conn>> conns_;
  std::vector<Waiter> waiters_;
};

class TCPStore : public Store {
 public:
  // is_master: also host the server (rank 0 semantics, ref: rendezvous.py:198 "start_daemon = rank == 0").
  // wait_for_workers: master blocks until world_size clients (incl. itself) have checked in.
  TCPStore(const std::string& host, int port, int world_size, bool is_master, Millis timeout,
           bool wait_for_workers = true);
  ~TCPStore() override;
  void set(const std::string& key, const Bytes& value) override;
  Bytes get(const std::string& key) override;
  int64_t add(const std::string& key, int64_t delta) override;
  Bytes compare_set(const std::string& key, const Bytes& expected, const Bytes& desired) override;
  using Store::wait;
  void wait(const std::vector<std::string>& keys, Millis timeout) override;
  bool check(const std::vector<std::string>& keys) override;
  bool delete_key(const std::string& key) override;
  int64_t num_keys() override;
  void append(const std::string& key, const Bytes& value) override;
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void queue_push(const std::string& key, const Bytes& value) override;
  Bytes queue_pop(const std::string& key, bool block) override;
  int64_t queue_len(const std::string& key) override;
  void ping();
  int port() const { return port_; }
  const std::string& host() const { return host_; }
  bool is_master() const { return server_ != nullptr; }

 private:
  std::vector<Bytes> call(uint8_t op, const std::vector<Bytes>& args, Millis timeout, uint8_t* status = nullptr);
  std::string host_;
  int port_;
  std::unique_ptr<TCPStoreServer> server_;
  std::mutex mu_;
  std::vector<Fd> idle_;  // pooled connections, one request in flight per connection
};

// Store-based barrier: every rank add()s a generation-scoped counter, last arrival releases.
void store_barrier(Store& store, const std::string& name, int rank, int world_size, Millis timeout);

}  // namespace pdt
