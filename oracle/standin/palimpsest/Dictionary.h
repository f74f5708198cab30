// SPDX-License-Identifier: Apache-2.0
//
// STAND-IN for <palimpsest/Dictionary.h> -- TEST INFRASTRUCTURE (see oracle/standin/Eigen/Core).
//
// The subset of palimpsest's nested dictionary the reference's observers and controllers use: a node is either a
// map of named children or a typed value; operator() descends (creating children on non-const access, throwing
// KeyError on const access), assignment stores a value, conversion / get<T>() read it back with its exact type.
#pragma once

#include <Eigen/Core>
#include <spdlog/spdlog.h>

#include <any>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>

#include "exceptions/KeyError.h"

namespace palimpsest {

class Dictionary {
 public:
  Dictionary() = default;
  Dictionary(const Dictionary&) = delete;
  Dictionary& operator=(const Dictionary&) = delete;

  bool has(const std::string& key) const { return children_.find(key) != children_.end(); }
  bool is_empty() const { return children_.empty() && !value_.has_value(); }

  Dictionary& operator()(const std::string& key) {
    auto it = children_.find(key);
    if (it == children_.end()) it = children_.emplace(key, std::make_unique<Dictionary>()).first;
    return *it->second;
  }
  const Dictionary& operator()(const std::string& key) const {
    auto it = children_.find(key);
    if (it == children_.end()) throw exceptions::KeyError(key);
    return *it->second;
  }

  std::vector<std::string> keys() const {
    std::vector<std::string> out;
    for (const auto& kv : children_) out.push_back(kv.first);
    return out;
  }

  template <typename T>
  Dictionary& operator=(const T& v) {
    value_ = v;
    return *this;
  }

  template <typename T>
  const T& as() const {
    if (!value_.has_value()) throw exceptions::TypeError("dictionary node holds no value");
    const T* p = std::any_cast<T>(&value_);
    if (!p) throw exceptions::TypeError(std::string("value is not a ") + typeid(T).name());
    return *p;
  }
  // implicit reads: `double x = dict("a")("b");`, `if (!dict("flag"))`, `double& v = action("velocity");`
  operator double() const { return as<double>(); }
  operator bool() const { return as<bool>(); }
  operator double&() {
    double* p = std::any_cast<double>(&value_);
    if (!p) throw exceptions::TypeError("value is not a double");
    return *p;
  }

  template <typename T>
  const T& get(const std::string& key) const {
    return (*this)(key).template as<T>();
  }
  template <typename T>
  T get(const std::string& key, const T& default_value) const {
    auto it = children_.find(key);
    if (it == children_.end()) return default_value;
    return it->second->template as<T>();
  }

  void clear() {
    children_.clear();
    value_.reset();
  }

 private:
  std::map<std::string, std::unique_ptr<Dictionary>> children_;
  std::any value_;
};

}  // namespace palimpsest
