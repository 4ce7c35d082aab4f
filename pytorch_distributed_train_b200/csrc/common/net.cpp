#include "net.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <thread>

namespace pdt {

std::string errno_str(const std::string& what) {
  return what + ": " + std::strerror(errno) + " (errno " + std::to_string(errno) + ")";
}

Fd& Fd::operator=(Fd&& o) noexcept {
  if (this != &o) {
    reset(o.fd_);
    o.fd_ = -1;
  }
  return *this;
}
Fd::~Fd() { reset(); }
void Fd::reset(int fd) {
  if (fd_ >= 0) ::close(fd_);
  fd_ = fd;
}

static int remaining_ms(Clock::time_point deadline) {
  auto left = std::chrono::duration_cast<Millis>(deadline - Clock::now()).count();
  if (left < 0) return 0;
  if (left > 1000 * 3600) return 1000 * 3600;
  return static_cast<int>(left);
}

void set_nodelay(int fd) {
  int one = 1;
  ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

void set_nonblocking(int fd, bool nb) {
  int fl = ::fcntl(fd, F_GETFL, 0);
  if (fl < 0) throw std::runtime_error(errno_str("fcntl(F_GETFL)"));
  fl = nb ? (fl | O_NONBLOCK) : (fl & ~O_NONBLOCK);
  if (::fcntl(fd, F_SETFL, fl) < 0) throw std::runtime_error(errno_str("fcntl(F_SETFL)"));
}

static struct addrinfo* resolve(const std::string& host, int port, bool passive) {
  struct addrinfo hints;
  std::memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_UNSPEC;
  hints.ai_socktype = SOCK_STREAM;
  if (passive) hints.ai_flags = AI_PASSIVE;
  struct addrinfo* res = nullptr;
  std::string p = std::to_string(port);
  const char* h = host.empty() ? nullptr : host.c_str();
  int rc = ::getaddrinfo(h, p.c_str(), &hints, &res);
  if (rc != 0) {
    throw std::runtime_error("getaddrinfo(" + host + ":" + p + "): " + gai_strerror(rc));
  }
  return res;
}

Fd tcp_listen(const std::string& host, int port, int* bound_port, int backlog) {
  struct addrinfo* res = resolve(host, port, true);
  std::string last_err = "no address";
  for (auto* ai = res; ai; ai = ai->ai_next) {
    Fd fd(::socket(ai->ai_family, ai->ai_socktype | SOCK_CLOEXEC, ai->ai_protocol));
    if (!fd.valid()) { last_err = errno_str("socket"); continue; }
    int one = 1;
    ::setsockopt(fd.get(), SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(fd.get(), ai->ai_addr, ai->ai_addrlen) != 0) { last_err = errno_str("bind"); continue; }
    if (::listen(fd.get(), backlog) != 0) { last_err = errno_str("listen"); continue; }
    if (bound_port) {
      struct sockaddr_storage ss;
      socklen_t len = sizeof(ss);
      ::getsockname(fd.get(), reinterpret_cast<struct sockaddr*>(&ss), &len);
      if (ss.ss_family == AF_INET)
        *bound_port = ntohs(reinterpret_cast<struct sockaddr_in*>(&ss)->sin_port);
      else
        *bound_port = ntohs(reinterpret_cast<struct sockaddr_in6*>(&ss)->sin6_port);
    }
    ::freeaddrinfo(res);
    return fd;
  }
  ::freeaddrinfo(res);
  throw std::runtime_error("tcp_listen(" + host + ":" + std::to_string(port) + ") failed: " + last_err);
}

Fd tcp_connect(const std::string& host, int port, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  std::string last_err;
  int backoff_ms = 2;
  while (true) {
    struct addrinfo* res = nullptr;
    try {
      res = resolve(host, port, false);
    } catch (const std::exception& e) {
      last_err = e.what();
    }
    for (auto* ai = res; ai; ai = ai->ai_next) {
      Fd fd(::socket(ai->ai_family, ai->ai_socktype | SOCK_CLOEXEC, ai->ai_protocol));
      if (!fd.valid()) { last_err = errno_str("socket"); continue; }
      if (::connect(fd.get(), ai->ai_addr, ai->ai_addrlen) == 0) {
        ::freeaddrinfo(res);
        set_nodelay(fd.get());
        return fd;
      }
      last_err = errno_str("connect");
    }
    if (res) ::freeaddrinfo(res);
    if (Clock::now() >= deadline) break;
    std::this_thread::sleep_for(Millis(backoff_ms));
    if (backoff_ms < 100) backoff_ms *= 2;
  }
  throw TimeoutError("tcp_connect(" + host + ":" + std::to_string(port) + ") timed out after " +
                     std::to_string(timeout.count()) + " ms; last error: " + last_err);
}

Fd tcp_accept(int listen_fd, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  while (true) {
    struct pollfd p = {listen_fd, POLLIN, 0};
    int rc = ::poll(&p, 1, remaining_ms(deadline));
    if (rc < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(errno_str("poll(accept)"));
    }
    if (rc == 0) throw TimeoutError("tcp_accept timed out");
    int fd = ::accept4(listen_fd, nullptr, nullptr, SOCK_CLOEXEC);
    if (fd < 0) {
      if (errno == EINTR || errno == EAGAIN || errno == ECONNABORTED) continue;
      throw std::runtime_error(errno_str("accept"));
    }
    set_nodelay(fd);
    return Fd(fd);
  }
}

void send_all(int fd, const void* buf, size_t n, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  const char* p = static_cast<const char*>(buf);
  size_t off = 0;
  while (off < n) {
    ssize_t w = ::send(fd, p + off, n - off, MSG_NOSIGNAL);
    if (w > 0) { off += static_cast<size_t>(w); continue; }
    if (w < 0 && (errno == EINTR)) continue;
    if (w < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
      struct pollfd pf = {fd, POLLOUT, 0};
      int rc = ::poll(&pf, 1, remaining_ms(deadline));
      if (rc == 0) throw TimeoutError("send timed out");
      continue;
    }
    if (w < 0 && (errno == EPIPE || errno == ECONNRESET)) throw PeerClosedError("peer closed connection during send");
    throw std::runtime_error(errno_str("send"));
  }
}

void recv_all(int fd, void* buf, size_t n, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  char* p = static_cast<char*>(buf);
  size_t off = 0;
  while (off < n) {
    struct pollfd pf = {fd, POLLIN, 0};
    int rc = ::poll(&pf, 1, remaining_ms(deadline));
    if (rc < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(errno_str("poll(recv)"));
    }
    if (rc == 0) throw TimeoutError("recv timed out after " + std::to_string(timeout.count()) + " ms");
    ssize_t r = ::recv(fd, p + off, n - off, 0);
    if (r > 0) { off += static_cast<size_t>(r); continue; }
    if (r == 0) throw PeerClosedError("peer closed connection during recv");
    if (errno == EINTR || errno == EAGAIN || errno == EWOULDBLOCK) continue;
    if (errno == ECONNRESET) throw PeerClosedError("connection reset by peer");
    throw std::runtime_error(errno_str("recv"));
  }
}

void send_recv(int send_fd, const void* sbuf, size_t sn, int recv_fd, void* rbuf, size_t rn, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  const char* sp = static_cast<const char*>(sbuf);
  char* rp = static_cast<char*>(rbuf);
  size_t soff = 0, roff = 0;
  while (soff < sn || roff < rn) {
    struct pollfd pf[2];
    int n = 0, si = -1, ri = -1;
    if (soff < sn) { pf[n] = {send_fd, POLLOUT, 0}; si = n++; }
    if (roff < rn) { pf[n] = {recv_fd, POLLIN, 0}; ri = n++; }
    int rc = ::poll(pf, n, remaining_ms(deadline));
    if (rc < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(errno_str("poll(send_recv)"));
    }
    if (rc == 0) throw TimeoutError("send_recv timed out after " + std::to_string(timeout.count()) + " ms");
    if (ri >= 0 && (pf[ri].revents & (POLLIN | POLLHUP | POLLERR))) {
      ssize_t r = ::recv(recv_fd, rp + roff, rn - roff, MSG_DONTWAIT);
      if (r > 0) roff += static_cast<size_t>(r);
      else if (r == 0) throw PeerClosedError("peer closed connection during send_recv");
      else if (!(errno == EINTR || errno == EAGAIN || errno == EWOULDBLOCK)) {
        if (errno == ECONNRESET) throw PeerClosedError("connection reset by peer");
        throw std::runtime_error(errno_str("recv"));
      }
    }
    if (si >= 0 && (pf[si].revents & (POLLOUT | POLLHUP | POLLERR))) {
      ssize_t w = ::send(send_fd, sp + soff, sn - soff, MSG_NOSIGNAL | MSG_DONTWAIT);
      if (w > 0) soff += static_cast<size_t>(w);
      else if (w < 0 && !(errno == EINTR || errno == EAGAIN || errno == EWOULDBLOCK)) {
        if (errno == EPIPE || errno == ECONNRESET) throw PeerClosedError("peer closed connection during send_recv");
        throw std::runtime_error(errno_str("send"));
      }
    }
  }
}

static socklen_t fill_abstract(struct sockaddr_un* addr, const std::string& name) {
  std::memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem entry to clean up
  size_t n = std::min(name.size(), sizeof(addr->sun_path) - 2);
  std::memcpy(addr->sun_path + 1, name.data(), n);
  return static_cast<socklen_t>(offsetof(struct sockaddr_un, sun_path) + 1 + n);
}

Fd unix_listen(const std::string& name) {
  Fd fd(::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0));
  if (!fd.valid()) throw std::runtime_error(errno_str("socket(AF_UNIX)"));
  struct sockaddr_un addr;
  socklen_t len = fill_abstract(&addr, name);
  if (::bind(fd.get(), reinterpret_cast<struct sockaddr*>(&addr), len) != 0)
    throw std::runtime_error(errno_str("bind(AF_UNIX " + name + ")"));
  if (::listen(fd.get(), 64) != 0) throw std::runtime_error(errno_str("listen(AF_UNIX)"));
  return fd;
}

Fd unix_connect(const std::string& name, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  struct sockaddr_un addr;
  socklen_t len = fill_abstract(&addr, name);
  while (true) {
    Fd fd(::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0));
    if (!fd.valid()) throw std::runtime_error(errno_str("socket(AF_UNIX)"));
    if (::connect(fd.get(), reinterpret_cast<struct sockaddr*>(&addr), len) == 0) return fd;
    if (Clock::now() >= deadline)
      throw TimeoutError("unix_connect(" + name + ") timed out: " + std::strerror(errno));
    std::this_thread::sleep_for(Millis(5));
  }
}

void send_fd(int sock, int fd_to_send, Millis timeout) {
  struct msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  char dummy = 'F';
  struct iovec iov = {&dummy, 1};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  alignas(struct cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd_to_send, sizeof(int));
  auto deadline = Clock::now() + timeout;
  while (true) {
    ssize_t w = ::sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (w == 1) return;
    if (w < 0 && (errno == EINTR || errno == EAGAIN)) {
      if (Clock::now() >= deadline) throw TimeoutError("send_fd timed out");
      continue;
    }
    throw std::runtime_error(errno_str("sendmsg(SCM_RIGHTS)"));
  }
}

int recv_fd(int sock, Millis timeout) {
  auto deadline = Clock::now() + timeout;
  while (true) {
    struct pollfd pf = {sock, POLLIN, 0};
    int rc = ::poll(&pf, 1, remaining_ms(deadline));
    if (rc < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(errno_str("poll(recv_fd)"));
    }
    if (rc == 0) throw TimeoutError("recv_fd timed out");
    struct msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    char dummy = 0;
    struct iovec iov = {&dummy, 1};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    alignas(struct cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    ssize_t r = ::recvmsg(sock, &msg, MSG_CMSG_CLOEXEC);
    if (r < 0) {
      if (errno == EINTR || errno == EAGAIN) continue;
      throw std::runtime_error(errno_str("recvmsg(SCM_RIGHTS)"));
    }
    if (r == 0) throw PeerClosedError("peer closed during recv_fd");
    for (struct cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c)) {
      if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
        int fd;
        std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
        return fd;
      }
    }
    throw std::runtime_error("recv_fd: message carried no SCM_RIGHTS payload");
  }
}

}  // namespace pdt
