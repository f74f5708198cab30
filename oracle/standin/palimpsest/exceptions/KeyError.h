// SPDX-License-Identifier: Apache-2.0
// STAND-IN for <palimpsest/exceptions/KeyError.h> -- TEST INFRASTRUCTURE.
#pragma once
#include <stdexcept>
#include <string>
namespace palimpsest::exceptions {
class KeyError : public std::runtime_error {
 public:
  explicit KeyError(const std::string& key) : std::runtime_error("key not found: " + key), key_(key) {}
  const std::string& key() const { return key_; }
 private:
  std::string key_;
};
class TypeError : public std::runtime_error {
 public:
  explicit TypeError(const std::string& what) : std::runtime_error(what) {}
};
}  // namespace palimpsest::exceptions
