// Small blocking/non-blocking TCP + AF_UNIX helpers shared by the store, the CPU
// collective backend and the fd-passing side channel.  Everything throws
// std::runtime_error with errno text on failure; timeouts are explicit.
#pragma once
#include <chrono>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace pdt {

using Clock = std::chrono::steady_clock;
using Millis = std::chrono::milliseconds;

struct TimeoutError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct PeerClosedError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// RAII fd.
class Fd {
 public:
  Fd() = default;
  explicit Fd(int fd) : fd_(fd) {}
  Fd(const Fd&) = delete;
  Fd& operator=(const Fd&) = delete;
  Fd(Fd&& o) noexcept : fd_(o.fd_) { o.fd_ = -1; }
  Fd& operator=(Fd&& o) noexcept;
  ~Fd();
  int get() const { return fd_; }
  int release() { int f = fd_; fd_ = -1; return f; }
  bool valid() const { return fd_ >= 0; }
  void reset(int fd = -1);

 private:
  int fd_ = -1;
};

// Listen on host:port (port 0 = ephemeral). Returns fd; *bound_port gets the real port.
Fd tcp_listen(const std::string& host, int port, int* bound_port, int backlog = 512);
// Connect with retry until `timeout` (connection refused is retried: the server
// may not be up yet — this is what makes "rank 0 hosts the store" race-free).
Fd tcp_connect(const std::string& host, int port, Millis timeout);
Fd tcp_accept(int listen_fd, Millis timeout);  // throws TimeoutError

void set_nodelay(int fd);
void set_nonblocking(int fd, bool nb);

// Blocking-with-deadline full send / full recv on a (blocking or non-blocking) fd.
void send_all(int fd, const void* buf, size_t n, Millis timeout);
void recv_all(int fd, void* buf, size_t n, Millis timeout);

// Full-duplex exchange: send `sn` bytes on send_fd while receiving `rn` bytes on recv_fd
// (they may be the same fd).  Never deadlocks on full socket buffers.
void send_recv(int send_fd, const void* sbuf, size_t sn, int recv_fd, void* rbuf, size_t rn,
               Millis timeout);

// AF_UNIX abstract-namespace datagram-free stream sockets for SCM_RIGHTS fd passing.
Fd unix_listen(const std::string& name);
Fd unix_connect(const std::string& name, Millis timeout);
void send_fd(int sock, int fd_to_send, Millis timeout);
int recv_fd(int sock, Millis timeout);

std::string errno_str(const std::string& what);

}  // namespace pdt
