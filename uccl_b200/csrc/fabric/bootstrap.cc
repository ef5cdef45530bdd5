#include "bootstrap.h"

#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <random>
#include <thread>

#include "../common/log.h"
#include "../common/param.h"

namespace ub {

UB_PARAM(BootstrapTimeoutSec, "BOOTSTRAP_TIMEOUT_SECS", 120)

namespace {

struct IdPayload {
  uint32_t magic;
  uint16_t port;
  uint16_t pad;
  uint64_t nonce;
  char ip[64];
};
static_assert(sizeof(IdPayload) <= sizeof(UniqueId), "id too big");
constexpr uint32_t kMagic = 0x55423230;  // "UB20"

bool write_full(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    c += w;
    n -= (size_t)w;
  }
  return true;
}
bool read_full(int fd, void* p, size_t n) {
  char* c = (char*)p;
  while (n) {
    ssize_t r = ::recv(fd, c, n, 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (r == 0) {
      errno = ECONNRESET;  // orderly close by the peer: report it as such, not with a stale errno
      return false;
    }
    c += r;
    n -= (size_t)r;
  }
  return true;
}

void relay_main(int lfd) {
  std::vector<int> conns;
  int nranks = -1;
  int accepted = 0;
  while (nranks < 0 || accepted < nranks) {
    int c = ::accept(lfd, nullptr, nullptr);
    if (c < 0) {
      if (errno == EINTR) continue;
      break;
    }
    int one = 1;
    setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    uint32_t hello[2];
    if (!read_full(c, hello, sizeof(hello))) {
      UB_WARN("bootstrap relay: a connection closed before its hello (%s)", strerror(errno));
      ::close(c);
      continue;
    }
    if (nranks < 0) {
      nranks = (int)hello[1];
      conns.assign(nranks, -1);
    }
    if ((int)hello[1] != nranks || (int)hello[0] >= nranks || conns[hello[0]] != -1) {
      UB_WARN("bootstrap relay: bad hello rank=%u nranks=%u", hello[0], hello[1]);
      ::close(c);
      continue;
    }
    conns[hello[0]] = c;
    ++accepted;
  }
  ::close(lfd);
  if (accepted != nranks) {
    UB_WARN("bootstrap relay: accept failed after %d of %d ranks (%s)", accepted, nranks, strerror(errno));
    for (int c : conns)
      if (c >= 0) ::close(c);
    return;
  }
  // rounds: gather one blob per rank, send the concatenation back to everyone
  std::vector<char> buf;
  bool alive = true;
  while (alive) {
    uint64_t bytes = 0;
    for (int r = 0; r < nranks && alive; ++r) {
      uint64_t b;
      if (!read_full(conns[r], &b, sizeof(b))) {
        // rank 0 hanging up between rounds is the normal end of a communicator; anything else is a rank that died
        // (or left) in the middle of an exchange: its peers will see their connection close
        if (r != 0) UB_WARN("bootstrap relay: rank %d of %d left in the middle of a round (%s)", r, nranks, strerror(errno));
        else UB_INFO(SUB_INIT, "bootstrap relay: rank 0 of %d hung up (%s): group ends", nranks, strerror(errno));
        alive = false;
        break;
      }
      if (r == 0) {
        bytes = b;
        buf.resize((size_t)bytes * nranks);
      } else if (b != bytes) {
        UB_WARN("bootstrap relay: size mismatch rank %d (%lu vs %lu)", r, (unsigned long)b, (unsigned long)bytes);
        alive = false;
        break;
      }
      if (bytes && !read_full(conns[r], buf.data() + (size_t)r * bytes, bytes)) alive = false;
    }
    if (!alive) break;
    for (int r = 0; r < nranks; ++r) {
      if (!buf.empty() && !write_full(conns[r], buf.data(), buf.size())) alive = false;
      if (buf.empty()) {
        char z = 0;
        if (!write_full(conns[r], &z, 1)) alive = false;
      }
    }
  }
  for (int c : conns)
    if (c >= 0) ::close(c);
}

std::string uds_name(uint64_t nonce, int rank) {
  char b[64];
  snprintf(b, sizeof(b), "ub_%016lx_%d", (unsigned long)nonce, rank);
  return std::string(b);
}

void fill_uds_addr(sockaddr_un& a, socklen_t& len, const std::string& name) {
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  a.sun_path[0] = 0;  // abstract namespace: no filesystem entry, vanishes with the process
  memcpy(a.sun_path + 1, name.data(), name.size());
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

}  // namespace

UniqueId Bootstrap::create_id(const char* ip_override) {
  UniqueId id;
  memset(&id, 0, sizeof(id));
  int lfd = ::socket(AF_INET, SOCK_STREAM, 0);
  UB_CHECK(lfd >= 0, "socket() failed: %s", strerror(errno));
  int one = 1;
  setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  std::string ip = ip_override ? std::string(ip_override) : param_load_str("BOOTSTRAP_IP", "127.0.0.1");
  sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_port = 0;
  UB_CHECK(inet_pton(AF_INET, ip.c_str(), &addr.sin_addr) == 1, "bad bootstrap ip %s", ip.c_str());
  UB_CHECK(::bind(lfd, (sockaddr*)&addr, sizeof(addr)) == 0, "bind(%s) failed: %s", ip.c_str(), strerror(errno));
  UB_CHECK(::listen(lfd, 64) == 0, "listen failed: %s", strerror(errno));
  socklen_t alen = sizeof(addr);
  getsockname(lfd, (sockaddr*)&addr, &alen);
  IdPayload p;
  memset(&p, 0, sizeof(p));
  p.magic = kMagic;
  p.port = ntohs(addr.sin_port);
  std::random_device rd;
  p.nonce = ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ ((uint64_t)getpid() << 16);
  snprintf(p.ip, sizeof(p.ip), "%s", ip.c_str());
  memcpy(id.data, &p, sizeof(p));
  std::thread(relay_main, lfd).detach();
  UB_INFO(SUB_INIT, "bootstrap root at %s:%u nonce=%016lx", p.ip, (unsigned)p.port, (unsigned long)p.nonce);
  return id;
}

Bootstrap::Bootstrap(const UniqueId& id, int rank, int nranks) : rank_(rank), nranks_(nranks) {
  IdPayload p;
  memcpy(&p, id.data, sizeof(p));
  UB_CHECK(p.magic == kMagic, "invalid unique id");
  UB_CHECK(rank >= 0 && rank < nranks, "bad rank %d/%d", rank, nranks);
  nonce_ = p.nonce;

  // fd-passing endpoint first, so that it exists before the first barrier returns
  uds_listen_ = ::socket(AF_UNIX, SOCK_STREAM, 0);
  UB_CHECK(uds_listen_ >= 0, "unix socket failed: %s", strerror(errno));
  sockaddr_un ua;
  socklen_t ulen;
  fill_uds_addr(ua, ulen, uds_name(nonce_, rank_));
  UB_CHECK(::bind(uds_listen_, (sockaddr*)&ua, ulen) == 0, "unix bind failed: %s", strerror(errno));
  UB_CHECK(::listen(uds_listen_, 64) == 0, "unix listen failed: %s", strerror(errno));

  sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_port = htons(p.port);
  inet_pton(AF_INET, p.ip, &addr.sin_addr);
  auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(ubParamBootstrapTimeoutSec());
  while (true) {
    sock_ = ::socket(AF_INET, SOCK_STREAM, 0);
    UB_CHECK(sock_ >= 0, "socket failed");
    if (::connect(sock_, (sockaddr*)&addr, sizeof(addr)) == 0) break;
    ::close(sock_);
    sock_ = -1;
    UB_CHECK(std::chrono::steady_clock::now() < deadline, "bootstrap connect to %s:%u timed out", p.ip,
             (unsigned)p.port);
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
  }
  int one = 1;
  setsockopt(sock_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  timeval tv;
  tv.tv_sec = ubParamBootstrapTimeoutSec();
  tv.tv_usec = 0;
  setsockopt(sock_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  uint32_t hello[2] = {(uint32_t)rank, (uint32_t)nranks};
  send_all(hello, sizeof(hello));
  barrier();
}

Bootstrap::~Bootstrap() {
  if (sock_ >= 0) ::close(sock_);
  if (uds_listen_ >= 0) ::close(uds_listen_);
}

void Bootstrap::send_all(const void* p, size_t n) {
  UB_CHECK(write_full(sock_, p, n), "bootstrap send failed: %s", strerror(errno));
}
void Bootstrap::recv_all(void* p, size_t n) {
  if (read_full(sock_, p, n)) return;
  const int e = errno;
  UB_CHECK(false, "bootstrap recv failed (group %016lx, rank %d/%d): %s", (unsigned long)nonce_, rank_, nranks_,
           e == EAGAIN || e == EWOULDBLOCK ? "timed out waiting for the other ranks (UCCL_B200_BOOTSTRAP_TIMEOUT_SECS)"
                                            : (e == ECONNRESET ? "a peer left the rendezvous" : strerror(e)));
}

void Bootstrap::allgather(const void* in, void* out, size_t bytes) {
  uint64_t b = bytes;
  send_all(&b, sizeof(b));
  if (bytes) send_all(in, bytes);
  if (bytes) {
    recv_all(out, bytes * (size_t)nranks_);
  } else {
    char z;
    recv_all(&z, 1);
  }
}

void Bootstrap::barrier() { allgather(nullptr, nullptr, 0); }

void Bootstrap::broadcast(void* buf, size_t bytes, int root) {
  std::vector<char> all(bytes * (size_t)nranks_);
  allgather(buf, all.data(), bytes);
  if (rank_ != root) memcpy(buf, all.data() + (size_t)root * bytes, bytes);
}

std::vector<std::vector<int>> Bootstrap::exchange_fds(const std::vector<int>& fds) {
  const int nf = (int)fds.size();
  UB_CHECK(nf > 0 && nf <= 64, "exchange_fds: bad fd count %d", nf);
  std::vector<std::vector<int>> result(nranks_, std::vector<int>(nf, -1));
  for (int i = 0; i < nf; ++i) result[rank_][i] = ::dup(fds[i]);
  barrier();
  // send to every peer (connect completes against the listen backlog; tiny message fits the socket buffer)
  for (int p = 0; p < nranks_; ++p) {
    if (p == rank_) continue;
    int s = ::socket(AF_UNIX, SOCK_STREAM, 0);
    UB_CHECK(s >= 0, "unix socket failed");
    sockaddr_un ua;
    socklen_t ulen;
    fill_uds_addr(ua, ulen, uds_name(nonce_, p));
    int tries = 0;
    while (::connect(s, (sockaddr*)&ua, ulen) != 0) {
      UB_CHECK(++tries < 2000, "unix connect to rank %d failed: %s", p, strerror(errno));
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    uint32_t hdr[3] = {(uint32_t)rank_, (uint32_t)nf, fd_round_};
    iovec iov{hdr, sizeof(hdr)};
    std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * nf));
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl.data();
    msg.msg_controllen = ctrl.size();
    cmsghdr* cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET;
    cm->cmsg_type = SCM_RIGHTS;
    cm->cmsg_len = CMSG_LEN(sizeof(int) * nf);
    memcpy(CMSG_DATA(cm), fds.data(), sizeof(int) * nf);
    UB_CHECK(::sendmsg(s, &msg, MSG_NOSIGNAL) == (ssize_t)sizeof(hdr), "sendmsg(SCM_RIGHTS) failed: %s",
             strerror(errno));
    ::close(s);
  }
  for (int k = 0; k < nranks_ - 1; ++k) {
    int c = ::accept(uds_listen_, nullptr, nullptr);
    UB_CHECK(c >= 0, "unix accept failed: %s", strerror(errno));
    uint32_t hdr[3];
    iovec iov{hdr, sizeof(hdr)};
    std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * nf));
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl.data();
    msg.msg_controllen = ctrl.size();
    ssize_t r = ::recvmsg(c, &msg, MSG_WAITALL);
    UB_CHECK(r == (ssize_t)sizeof(hdr), "recvmsg failed: %s", strerror(errno));
    cmsghdr* cm = CMSG_FIRSTHDR(&msg);
    UB_CHECK(cm && cm->cmsg_type == SCM_RIGHTS && hdr[1] == (uint32_t)nf && hdr[2] == fd_round_ &&
                 hdr[0] < (uint32_t)nranks_,
             "bad fd message from peer");
    memcpy(result[hdr[0]].data(), CMSG_DATA(cm), sizeof(int) * nf);
    ::close(c);
  }
  ++fd_round_;
  barrier();
  return result;
}

}  // namespace ub
