// Leveled, subsystem-filtered logger + CHECK macros.
// Behavioural parity with the reference's UCCL_LOG family (include/util/debug.h:1-60,302,365-400):
// env UCCL_DEBUG={FATAL,ERROR,WARN,INFO,TRACE}, UCCL_DEBUG_SUBSYS=comma list (^ inverts),
// hostname:pid:tid prefix, *_EVERY_N / *_FIRST_N / *_IF helpers, CHECK/DCHECK that abort.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <sys/syscall.h>
#include <unistd.h>

namespace ub {

enum LogLevel : int { LOG_FATAL = 0, LOG_ERROR = 1, LOG_WARN = 2, LOG_INFO = 3, LOG_TRACE = 4 };
enum LogSubsys : uint32_t {
  SUB_INIT = 1u << 0,
  SUB_FABRIC = 1u << 1,
  SUB_COLL = 1u << 2,
  SUB_EP = 1u << 3,
  SUB_P2P = 1u << 4,
  SUB_UTIL = 1u << 5,
  SUB_NCCL = 1u << 6,
  SUB_CC = 1u << 7,
  SUB_NET = 1u << 8,
  SUB_ALL = 0xffffffffu
};

struct LogState {
  int level = LOG_WARN;
  uint32_t subsys = SUB_ALL;
  FILE* out = stderr;
  char host[64] = {0};
  std::mutex mu;
};

inline LogState& log_state() {
  static LogState* st = [] {
    auto* s = new LogState();
    const char* lv = getenv("UCCL_B200_DEBUG");
    if (!lv) lv = getenv("UCCL_DEBUG");
    if (lv) {
      if (!strcasecmp(lv, "FATAL")) s->level = LOG_FATAL;
      else if (!strcasecmp(lv, "ERROR")) s->level = LOG_ERROR;
      else if (!strcasecmp(lv, "WARN")) s->level = LOG_WARN;
      else if (!strcasecmp(lv, "INFO")) s->level = LOG_INFO;
      else if (!strcasecmp(lv, "TRACE") || !strcasecmp(lv, "VERSION")) s->level = LOG_TRACE;
    }
    const char* ss = getenv("UCCL_B200_DEBUG_SUBSYS");
    if (!ss) ss = getenv("UCCL_DEBUG_SUBSYS");
    if (ss && *ss) {
      bool invert = (*ss == '^');
      if (invert) ++ss;
      uint32_t mask = 0;
      std::string str(ss);
      size_t pos = 0;
      while (pos <= str.size()) {
        size_t e = str.find(',', pos);
        if (e == std::string::npos) e = str.size();
        std::string tok = str.substr(pos, e - pos);
        if (!strcasecmp(tok.c_str(), "INIT")) mask |= SUB_INIT;
        else if (!strcasecmp(tok.c_str(), "FABRIC")) mask |= SUB_FABRIC;
        else if (!strcasecmp(tok.c_str(), "COLL")) mask |= SUB_COLL;
        else if (!strcasecmp(tok.c_str(), "EP")) mask |= SUB_EP;
        else if (!strcasecmp(tok.c_str(), "P2P")) mask |= SUB_P2P;
        else if (!strcasecmp(tok.c_str(), "UTIL")) mask |= SUB_UTIL;
        else if (!strcasecmp(tok.c_str(), "NCCL")) mask |= SUB_NCCL;
        else if (!strcasecmp(tok.c_str(), "CC")) mask |= SUB_CC;
        else if (!strcasecmp(tok.c_str(), "NET")) mask |= SUB_NET;
        else if (!strcasecmp(tok.c_str(), "ALL")) mask |= SUB_ALL;
        pos = e + 1;
      }
      s->subsys = invert ? ~mask : mask;
    }
    const char* f = getenv("UCCL_B200_DEBUG_FILE");
    if (f && *f) {
      FILE* fp = fopen(f, "a");
      if (fp) s->out = fp;
    }
    gethostname(s->host, sizeof(s->host) - 1);
    return s;
  }();
  return *st;
}

inline void set_log_level(int lv) { log_state().level = lv; }

inline void log_emit(int level, uint32_t subsys, const char* file, int line, const char* fmt, ...) {
  LogState& s = log_state();
  if (level > s.level) return;
  if (level > LOG_WARN && !(s.subsys & subsys)) return;
  static const char* names[] = {"FATAL", "ERROR", "WARN", "INFO", "TRACE"};
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  const char* base = strrchr(file, '/');
  base = base ? base + 1 : file;
  std::lock_guard<std::mutex> g(s.mu);
  fprintf(s.out, "%s:%d:%ld [uccl_b200 %s] %s:%d %s\n", s.host, (int)getpid(), (long)syscall(SYS_gettid),
          names[level], base, line, buf);
  fflush(s.out);
}

}  // namespace ub

#define UB_LOG(level, subsys, ...) ::ub::log_emit(::ub::level, ::ub::subsys, __FILE__, __LINE__, __VA_ARGS__)
#define UB_INFO(subsys, ...) UB_LOG(LOG_INFO, subsys, __VA_ARGS__)
#define UB_WARN(...) UB_LOG(LOG_WARN, SUB_ALL, __VA_ARGS__)
#define UB_ERROR(...) UB_LOG(LOG_ERROR, SUB_ALL, __VA_ARGS__)
#define UB_TRACE(subsys, ...) UB_LOG(LOG_TRACE, subsys, __VA_ARGS__)
#define UB_LOG_IF(cond, level, subsys, ...) \
  do {                                      \
    if (cond) UB_LOG(level, subsys, __VA_ARGS__); \
  } while (0)
#define UB_LOG_EVERY_N(n, level, subsys, ...)                                   \
  do {                                                                          \
    static std::atomic<uint64_t> _ub_cnt{0};                                    \
    if ((_ub_cnt.fetch_add(1, std::memory_order_relaxed) % (n)) == 0) UB_LOG(level, subsys, __VA_ARGS__); \
  } while (0)
#define UB_LOG_FIRST_N(n, level, subsys, ...)                                   \
  do {                                                                          \
    static std::atomic<uint64_t> _ub_cnt{0};                                    \
    if (_ub_cnt.fetch_add(1, std::memory_order_relaxed) < (uint64_t)(n)) UB_LOG(level, subsys, __VA_ARGS__); \
  } while (0)

// Errors surface as C++ exceptions (translated to Python RuntimeError by pybind11).
#define UB_THROW(...)                                              \
  do {                                                             \
    char _ub_buf[1024];                                            \
    snprintf(_ub_buf, sizeof(_ub_buf), __VA_ARGS__);               \
    UB_ERROR("%s", _ub_buf);                                       \
    throw std::runtime_error(std::string("uccl_b200: ") + _ub_buf); \
  } while (0)
#define UB_CHECK(cond, ...)                       \
  do {                                            \
    if (!(cond)) {                                \
      UB_ERROR("CHECK failed: %s", #cond);        \
      UB_THROW(__VA_ARGS__);                      \
    }                                             \
  } while (0)
#ifndef NDEBUG
#define UB_DCHECK(cond, ...) UB_CHECK(cond, __VA_ARGS__)
#else
#define UB_DCHECK(cond, ...) \
  do {                       \
  } while (0)
#endif
