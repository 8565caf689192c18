// hammlet_b200 host side — common types and small helpers.
//
// The host side mirrors the reference's C++ model surface (class and method names, argument meaning,
// error behaviour) so that code written against the reference's headers reads the same here; the
// per-sweep bodies call the C ABI in include/hammlet_b200.h instead of walking blocks on the CPU.
// Reference: src/includes.hpp (real_t :10, marginal_t :13), src/Distribution.hpp:15 (rng_t).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef HAMMLET_REAL
#define HAMMLET_REAL float  // the reference hard-codes float; -DHAMMLET_REAL=double builds the fp64-host variant
#endif
typedef HAMMLET_REAL real_t;
typedef int16_t marginal_t;
typedef std::mt19937 rng_t;

namespace hammlet {

const real_t inf = std::numeric_limits<real_t>::infinity();

inline bool fileExists(const std::string& path) {
  std::ifstream f(path.c_str());
  return f.good();
}

// "a<sep>b<sep>c<finalSep>" via operator<< (reference: utils.hpp:101-124, without the padding arguments)
template <typename T>
std::string concat(const std::vector<T>& v, const std::string& sep = "\t", const std::string& finalSep = "") {
  std::ostringstream os;
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) os << sep;
    os << v[i];
  }
  os << finalSep;
  return os.str();
}

}  // namespace hammlet
