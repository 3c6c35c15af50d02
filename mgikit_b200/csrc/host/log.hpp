// Minimal stdout logger in the spirit of the reference's `log` + `env_logger`
// (/root/reference/src/main.rs:713-733): level from -l/--log-level.
#pragma once

#include <string>

namespace mgikit {

enum LogLevel { LOG_ERROR = 0, LOG_WARN = 1, LOG_INFO = 2, LOG_DEBUG = 3, LOG_TRACE = 4 };

void set_log_level(const std::string& name);
void log_msg(LogLevel lvl, const std::string& msg);
inline void log_error(const std::string& m) { log_msg(LOG_ERROR, m); }
inline void log_warn(const std::string& m) { log_msg(LOG_WARN, m); }
inline void log_info(const std::string& m) { log_msg(LOG_INFO, m); }
inline void log_debug(const std::string& m) { log_msg(LOG_DEBUG, m); }

}  // namespace mgikit
