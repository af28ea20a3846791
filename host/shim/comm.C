#include "comm.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace {
int envInt(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}
[[noreturn]] void fail(const std::string &what) { throw std::runtime_error("Comm: " + what + " (" + std::strerror(errno) + ")"); }
}  // namespace

Comm &Comm::world() {
  static Comm c;
  return c;
}

Comm::Comm() {
  _size = envInt("WORLD_SIZE", 1);
  _rank = envInt("RANK", 0);
  _local_rank = envInt("LOCAL_RANK", _rank);
  if (_size <= 1) {
    _size = 1;
    _rank = 0;
    return;
  }
  if (_rank < 0 || _rank >= _size) throw std::runtime_error("Comm: RANK must be in [0, WORLD_SIZE)");
  const char *addr = std::getenv("MASTER_ADDR");
  const std::string host = addr && *addr ? addr : "127.0.0.1";
  // MRL_COMM_PORT: torchrun keeps MASTER_PORT for its own store when it launches the ranks itself
  const int port = envInt("MRL_COMM_PORT", envInt("MASTER_PORT", 29533) + (std::getenv("TORCHELASTIC_RUN_ID") ? 1 : 0));
  sockaddr_in sa;
  std::memset(&sa, 0, sizeof sa);
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)port);
  const int one = 1;
  if (_rank == 0) {
    const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) fail("socket");
    ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    if (::bind(ls, (sockaddr *)&sa, sizeof sa) < 0) fail("bind to port " + std::to_string(port));
    if (::listen(ls, _size) < 0) fail("listen");
    _peers.assign(_size, -1);
    // a rank that never shows up is an error after MRL_COMM_TIMEOUT seconds (default 120), not a hang
    const int timeout_ms = envInt("MRL_COMM_TIMEOUT", 120) * 1000;
    for (int i = 1; i < _size; ++i) {
      pollfd pfd{ls, POLLIN, 0};
      const int ready = ::poll(&pfd, 1, timeout_ms);
      if (ready == 0)
        throw std::runtime_error("Comm: only " + std::to_string(i) + " of " + std::to_string(_size) + " ranks reached the rendezvous on port " +
                                 std::to_string(port) + " within " + std::to_string(timeout_ms / 1000) + " s");
      if (ready < 0) fail("poll");
      const int fd = ::accept(ls, nullptr, nullptr);
      if (fd < 0) fail("accept");
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      int32_t r = -1;
      recvAll(fd, &r, sizeof r);
      if (r < 1 || r >= _size || _peers[r] >= 0) throw std::runtime_error("Comm: unexpected rank " + std::to_string(r) + " at the rendezvous");
      _peers[r] = fd;
    }
    ::close(ls);
  } else {
    if (::inet_pton(AF_INET, host.c_str(), &sa.sin_addr) != 1) throw std::runtime_error("Comm: MASTER_ADDR must be an IPv4 address, got '" + host + "'");
    for (int attempt = 0;; ++attempt) {
      _hub = ::socket(AF_INET, SOCK_STREAM, 0);
      if (_hub < 0) fail("socket");
      if (::connect(_hub, (sockaddr *)&sa, sizeof sa) == 0) break;
      ::close(_hub);
      if (attempt > 600) fail("connect to " + host + ":" + std::to_string(port));
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    ::setsockopt(_hub, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    const int32_t r = _rank;
    sendAll(_hub, &r, sizeof r);
  }
}

Comm::~Comm() {
  if (_hub >= 0) ::close(_hub);
  for (int fd : _peers)
    if (fd >= 0) ::close(fd);
}

void Comm::sendAll(int fd, const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) {
    const ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
    if (k <= 0) fail("send (a peer process has gone away)");
    c += k;
    n -= (size_t)k;
  }
}
void Comm::recvAll(int fd, void *p, size_t n) {
  char *c = (char *)p;
  while (n) {
    const ssize_t k = ::recv(fd, c, n, 0);
    if (k <= 0) fail("recv (a peer process has gone away)");
    c += k;
    n -= (size_t)k;
  }
}

void Comm::allgather(const void *in, size_t bytes, void *out) {
  char *o = (char *)out;
  if (_size == 1) {
    std::memcpy(o, in, bytes);
    return;
  }
  if (_rank == 0) {
    std::memcpy(o, in, bytes);
    for (int r = 1; r < _size; ++r) recvAll(_peers[r], o + (size_t)r * bytes, bytes);
    for (int r = 1; r < _size; ++r) sendAll(_peers[r], o, bytes * _size);
  } else {
    sendAll(_hub, in, bytes);
    recvAll(_hub, o, bytes * _size);
  }
}

void Comm::allreduce(double *v, size_t n, Op op) {
  if (_size == 1 || n == 0) return;
  std::vector<double> all(n * _size);
  allgather(v, n * sizeof(double), all.data());
  for (size_t i = 0; i < n; ++i) {
    double a = all[i];  // rank order: the same summation order on every rank
    for (int r = 1; r < _size; ++r) {
      const double b = all[(size_t)r * n + i];
      a = op == SUM ? a + b : op == MIN ? (b < a ? b : a) : (b > a ? b : a);
    }
    v[i] = a;
  }
}

void Comm::barrier() {
  double z = 0;
  allreduce(&z, 1, SUM);
}
