#include "log.hpp"

#include <cstdio>
#include <ctime>
#include <mutex>

namespace mgikit {

static int g_level = LOG_INFO;
static std::mutex g_mu;

void set_log_level(const std::string& name) {
  if (name == "error") g_level = LOG_ERROR;
  else if (name == "warn") g_level = LOG_WARN;
  else if (name == "info") g_level = LOG_INFO;
  else if (name == "debug") g_level = LOG_DEBUG;
  else if (name == "trace") g_level = LOG_TRACE;
  else if (name == "off") g_level = -1;
}

void log_msg(LogLevel lvl, const std::string& msg) {
  if ((int)lvl > g_level) return;
  static const char* names[] = {"ERROR", "WARN ", "INFO ", "DEBUG", "TRACE"};
  char ts[32];
  time_t t = time(nullptr);
  struct tm tmv;
  gmtime_r(&t, &tmv);
  strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &tmv);
  std::lock_guard<std::mutex> lk(g_mu);
  std::printf("[%s %s mgikit] %s\n", ts, names[lvl], msg.c_str());
  std::fflush(stdout);
}

}  // namespace mgikit
