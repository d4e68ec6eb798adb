// STAND-IN for cnr_logger (CNR-STIIMA-IRAS/cnr_logger, absent from this image): a logger that drops
// everything.  Test infrastructure for oracle/_ref only.
#pragma once
#include <memory>
#include <sstream>
#include <string>
namespace cnr_logger
{
class TraceLogger
{
public:
  TraceLogger() {}
  TraceLogger(const std::string&, const std::string& = "", bool = true, bool = true) {}
};
typedef std::shared_ptr<TraceLogger> TraceLoggerPtr;
}  // namespace cnr_logger
// the message expression is still compiled (so the reference's operator<< uses are type-checked) but never run
#define CNR_SHIM_LOG(logger, ...)                                                                                      \
  do                                                                                                                   \
  {                                                                                                                    \
    if (false)                                                                                                         \
    {                                                                                                                  \
      (void)(logger);                                                                                                  \
      std::stringstream cnr_shim_ss;                                                                                   \
      cnr_shim_ss << __VA_ARGS__;                                                                                      \
    }                                                                                                                  \
  } while (0)
#define CNR_TRACE(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
#define CNR_DEBUG(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
#define CNR_INFO(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
#define CNR_WARN(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
#define CNR_ERROR(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
#define CNR_FATAL(logger, ...) CNR_SHIM_LOG(logger, __VA_ARGS__)
