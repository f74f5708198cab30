// SPDX-License-Identifier: Apache-2.0
// STAND-IN for <spdlog/spdlog.h> -- TEST INFRASTRUCTURE (see oracle/standin/Eigen/Core). Logging calls are no-ops.
#pragma once
namespace spdlog {
template <typename... Args> inline void debug(const Args&...) {}
template <typename... Args> inline void info(const Args&...) {}
template <typename... Args> inline void warn(const Args&...) {}
template <typename... Args> inline void error(const Args&...) {}
}  // namespace spdlog
