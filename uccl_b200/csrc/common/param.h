// Cached env-var parameters with layered config files.
// Behavioural parity with UCCL_PARAM (collective/rdma/param.h:16-43, param.cc:46-93):
//   UB_PARAM(Name, "ENV_SUFFIX", default) defines `int64_t ubParamName()` that reads
//   UCCL_B200_<ENV_SUFFIX> (falling back to UCCL_<ENV_SUFFIX>) once and caches it.
//   Config files ($UCCL_B200_CONF_FILE, ~/.uccl_b200.conf, /etc/uccl_b200.conf) hold
//   KEY=VALUE lines that are injected into the environment (without overriding it).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

namespace ub {

inline void param_load_conf_file(const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) return;
  char line[1024];
  while (fgets(line, sizeof(line), f)) {
    char* p = line;
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == '#' || *p == '\n' || *p == 0) continue;
    char* eq = strchr(p, '=');
    if (!eq) continue;
    *eq = 0;
    char* v = eq + 1;
    size_t n = strlen(v);
    while (n && (v[n - 1] == '\n' || v[n - 1] == '\r' || v[n - 1] == ' ')) v[--n] = 0;
    char* ke = eq;
    while (ke > p && (ke[-1] == ' ' || ke[-1] == '\t')) *--ke = 0;
    setenv(p, v, /*overwrite=*/0);
  }
  fclose(f);
}

inline void param_init_env() {
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char* c = getenv("UCCL_B200_CONF_FILE")) param_load_conf_file(c);
    if (const char* h = getenv("HOME")) {
      std::string p = std::string(h) + "/.uccl_b200.conf";
      param_load_conf_file(p.c_str());
    }
    param_load_conf_file("/etc/uccl_b200.conf");
  });
}

inline int64_t param_load(const char* suffix, int64_t dflt) {
  param_init_env();
  std::string k1 = std::string("UCCL_B200_") + suffix;
  std::string k2 = std::string("UCCL_") + suffix;
  const char* v = getenv(k1.c_str());
  if (!v) v = getenv(k2.c_str());
  if (!v || !*v) return dflt;
  char* end = nullptr;
  long long x = strtoll(v, &end, 0);
  if (end == v) return dflt;
  if (end && (*end == 'k' || *end == 'K')) x <<= 10;
  if (end && (*end == 'm' || *end == 'M')) x <<= 20;
  if (end && (*end == 'g' || *end == 'G')) x <<= 30;
  return (int64_t)x;
}

inline std::string param_load_str(const char* suffix, const char* dflt) {
  param_init_env();
  std::string k1 = std::string("UCCL_B200_") + suffix;
  std::string k2 = std::string("UCCL_") + suffix;
  const char* v = getenv(k1.c_str());
  if (!v) v = getenv(k2.c_str());
  return v ? std::string(v) : std::string(dflt);
}

}  // namespace ub

#define UB_PARAM(name, env, dflt)                                         \
  static inline int64_t ubParam##name() {                                 \
    static std::atomic<int64_t> cached{INT64_MIN};                        \
    int64_t v = cached.load(std::memory_order_relaxed);                   \
    if (v == INT64_MIN) {                                                 \
      v = ::ub::param_load(env, dflt);                                    \
      cached.store(v, std::memory_order_relaxed);                         \
    }                                                                     \
    return v;                                                             \
  }
