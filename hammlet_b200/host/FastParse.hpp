// hammlet_b200 host side — input pipeline: whitespace-separated numbers -> float, many times faster than the
// `input >> v` loop of the reference (wavelet.hpp:131 spends about 0.5 us per value in num_get) and with the same
// result for every input:
//   * a number is what std::num_get accumulates: [+-] digits [. digits] [(e|E) [+-] digits]; no hex, inf, nan
//   * it is converted with correct rounding (std::from_chars; num_get converts with strtof, also correctly rounded)
//   * reading stops at the first token that num_get would reject (junk, incomplete exponent, overflow to
//     +-HUGE_VALF); the values before it are kept, exactly as `while (input >> v)` leaves them
// The text is split at whitespace into pieces parsed by several threads; the pieces are joined in order and cut
// at the first failure.
#pragma once
#include <charconv>
#include <cmath>
#include <cstring>
#include <fstream>
#include <istream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace fastparse {

inline bool isSpace(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
inline bool isDigit(char c) { return c >= '0' && c <= '9'; }

// Parses numbers of [p, end) into out.  Returns false if a token was rejected (reading must stop there).
inline bool parsePiece(const char* p, const char* end, std::vector<float>& out) {
  while (true) {
    while (p < end && isSpace(*p)) ++p;
    if (p >= end) return true;
    // the characters std::num_get::_M_extract_float would accumulate
    const char* q = p;
    if (q < end && (*q == '+' || *q == '-')) ++q;
    const char* digits0 = q;
    while (q < end && isDigit(*q)) ++q;
    bool anyDigit = q > digits0;
    if (q < end && *q == '.') {
      ++q;
      const char* f0 = q;
      while (q < end && isDigit(*q)) ++q;
      anyDigit = anyDigit || q > f0;
    }
    if (!anyDigit) return false;
    if (q < end && (*q == 'e' || *q == 'E')) {
      const char* e = q + 1;
      if (e < end && (*e == '+' || *e == '-')) ++e;
      const char* e0 = e;
      while (e < end && isDigit(*e)) ++e;
      if (e == e0) return false;  // "1e", "1e+": strtof leaves characters behind and num_get sets failbit
      q = e;
    }
    const char* first = (*p == '+') ? p + 1 : p;  // from_chars takes no leading plus sign
    float v = 0.f;
    const std::from_chars_result r = std::from_chars(first, q, v);
    if (r.ec == std::errc::result_out_of_range) {
      // strtof: overflow gives +-HUGE_VALF, which num_get reports as failure; underflow gives the rounded
      // (possibly zero or subnormal) value, which it accepts
      const std::string tok(p, q);
      v = std::strtof(tok.c_str(), nullptr);
      if (std::isinf(v)) return false;
    } else if (r.ec != std::errc() || r.ptr != q) {
      return false;
    }
    out.push_back(v);
    p = q;  // the next extraction starts right here: "1.5abc" yields 1.5 and then fails, "1..2" yields 1 and 0.2
  }
}

inline void parseFloats(const char* data, size_t n, std::vector<float>& out, unsigned threads = 0) {
  if (threads == 0) {
    threads = std::thread::hardware_concurrency();
    if (threads == 0) threads = 1;
    if (threads > 32) threads = 32;
  }
  if (n < (1u << 20)) threads = 1;
  std::vector<size_t> cut(threads + 1, n);
  cut[0] = 0;
  for (unsigned t = 1; t < threads; ++t) {
    size_t c = n / threads * t;
    if (c < cut[t - 1]) c = cut[t - 1];
    while (c < n && !isSpace(data[c])) ++c;  // pieces begin at whitespace
    cut[t] = c;
  }
  std::vector<std::vector<float>> parts(threads);
  std::vector<char> ok(threads, 1);
  auto work = [&](unsigned t) {
    parts[t].reserve((cut[t + 1] - cut[t]) / 6 + 16);
    ok[t] = parsePiece(data + cut[t], data + cut[t + 1], parts[t]) ? 1 : 0;
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  size_t total = 0;
  for (unsigned t = 0; t < threads; ++t) {
    total += parts[t].size();
    if (!ok[t]) break;
  }
  out.reserve(out.size() + total);
  for (unsigned t = 0; t < threads; ++t) {
    out.insert(out.end(), parts[t].begin(), parts[t].end());
    if (!ok[t]) break;
  }
}

// whole stream / file into memory
inline std::string slurp(std::istream& in) {
  std::string s;
  char buf[1 << 16];
  while (in.read(buf, sizeof(buf)) || in.gcount() > 0) s.append(buf, (size_t)in.gcount());
  return s;
}

}  // namespace fastparse
