// SPDX-License-Identifier: Apache-2.0
// STAND-IN for <spdlog/fmt/ostr.h> -- TEST INFRASTRUCTURE.
#pragma once
#include "../spdlog.h"
