// hammlet_b200 host side — input pipeline: whitespace-separated numbers -> float, many times faster than the
// `input >> v` loop of the reference (wavelet.hpp:131 spends about 0.5 us per value in num_get) and with the same
// result for every input:
//   * a number is what std::num_get accumulates: [+-] digits [. digits] [(e|E) [+-] digits]; no hex, inf, nan
//   * it is converted with correct rounding (std::from_chars; num_get converts with strtof, also correctly rounded)
//   * reading stops at the first token that num_get would reject (junk, incomplete exponent, overflow to
//     +-HUGE_VALF); the values before it are kept, exactly as `while (input >> v)` leaves them
// The text is split at whitespace into pieces parsed by several threads; the pieces are joined in order and cut
// at the first failure.
//
// Besides plain text the pipeline reads (readValues): gzip'd text — what the reference's preprocessing writes,
// bin/samToCounts: `*-count.csv.gz`, one count per line, which the reference itself can only take through `zcat |` —
// inflated piece by piece on one thread while the finished pieces are parsed on the others; and raw little-endian
// float32, for inputs that have been through the parser once (1e9 observations are 9 GB of text but 4 GB of floats).
#pragma once
#include <zlib.h>

#include <charconv>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <fstream>
#include <future>
#include <istream>
#include <iterator>
#include <memory>
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

enum class Format { Auto, Text, Gzip, F32 };
inline Format formatFromName(const std::string& name) {
  if (name == "auto") return Format::Auto;
  if (name == "text") return Format::Text;
  if (name == "gz" || name == "gzip") return Format::Gzip;
  if (name == "f32" || name == "float32") return Format::F32;
  throw std::runtime_error("Unknown input format " + name + " (auto, text, gz, f32)!");
}
inline bool isGzip(const char* data, size_t n) { return n >= 2 && (unsigned char)data[0] == 0x1f && (unsigned char)data[1] == 0x8b; }

// gzip'd text (concatenated members included): one thread inflates into pieces of ~8 MB cut at whitespace, the pieces
// are parsed as they become available; the values are joined in order and cut at the first rejected token.
inline void parseGzip(const char* data, size_t n, std::vector<float>& out, unsigned threads = 0) {
  if (threads == 0) {
    threads = std::thread::hardware_concurrency();
    if (threads == 0) threads = 1;
    if (threads > 32) threads = 32;
  }
  struct Piece {
    std::string text;
    std::vector<float> values;
    bool ok = true;
  };
  std::vector<std::unique_ptr<Piece>> pieces;
  std::vector<std::future<void>> jobs;
  auto submit = [&](std::string&& text) {
    pieces.emplace_back(new Piece());
    Piece* pc = pieces.back().get();
    pc->text = std::move(text);
    if (jobs.size() >= threads) jobs[jobs.size() - threads].wait();  // bounded look-ahead
    jobs.push_back(std::async(std::launch::async, [pc]() {
      pc->values.reserve(pc->text.size() / 6 + 16);
      pc->ok = parsePiece(pc->text.data(), pc->text.data() + pc->text.size(), pc->values);
      std::string().swap(pc->text);
    }));
  };
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) throw std::runtime_error("Cannot initialise zlib!");
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(data));
  size_t left = n;
  const size_t kPiece = 8u << 20;
  std::string cur;
  cur.reserve(kPiece + (1u << 16));
  std::vector<char> buf(1u << 18);
  int rc = Z_OK;
  while (true) {
    if (zs.avail_in == 0 && left > 0) {
      const size_t take = left > (1u << 30) ? (1u << 30) : left;
      zs.avail_in = (uInt)take;
      left -= take;
    }
    zs.next_out = reinterpret_cast<Bytef*>(buf.data());
    zs.avail_out = (uInt)buf.size();
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) {
      inflateEnd(&zs);
      throw std::runtime_error("Cannot decompress the input (corrupt gzip data)!");
    }
    cur.append(buf.data(), buf.size() - zs.avail_out);
    if (cur.size() >= kPiece) {  // hand over everything up to the last whitespace
      size_t c = cur.size();
      while (c > 0 && !isSpace(cur[c - 1])) --c;
      if (c > 0) {
        std::string rest = cur.substr(c);
        cur.resize(c);
        submit(std::move(cur));
        cur = std::move(rest);
        cur.reserve(kPiece + (1u << 16));
      }
    }
    if (rc == Z_STREAM_END) {
      if (zs.avail_in == 0 && left == 0) break;
      if (inflateReset(&zs) != Z_OK) break;  // next member of a multi-member file
    } else if (zs.avail_in == 0 && left == 0 && zs.avail_out != 0) {
      inflateEnd(&zs);
      throw std::runtime_error("Cannot decompress the input (truncated gzip data)!");
    }
  }
  inflateEnd(&zs);
  if (!cur.empty()) submit(std::move(cur));
  for (auto& j : jobs) j.wait();
  size_t total = 0;
  for (auto& pc : pieces) {
    total += pc->values.size();
    if (!pc->ok) break;
  }
  out.reserve(out.size() + total);
  for (auto& pc : pieces) {
    out.insert(out.end(), pc->values.begin(), pc->values.end());
    if (!pc->ok) break;
  }
}

// The values of an input in any of the supported formats.  Auto: gzip by its magic number, else text.
inline void parseAny(const char* data, size_t n, Format format, std::vector<float>& out, unsigned threads = 0) {
  if (format == Format::Auto) format = isGzip(data, n) ? Format::Gzip : Format::Text;
  if (format == Format::Gzip) {
    if (!isGzip(data, n)) throw std::runtime_error("Input is not in gzip format!");
    parseGzip(data, n, out, threads);
  } else if (format == Format::F32) {
    if (n % sizeof(float) != 0) throw std::runtime_error("Raw float32 input must hold a whole number of 4-byte values!");
    const size_t count = n / sizeof(float), at = out.size();
    out.resize(at + count);
    std::memcpy(out.data() + at, data, n);  // little-endian hosts only (x86-64, aarch64)
  } else {
    parseFloats(data, n, out, threads);
  }
}

// whole file into memory with one read (the istream loop above moves 64 kB at a time)
inline std::string slurpFile(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("Cannot read from input file " + path + "!");
  std::string s;
  if (std::fseek(f, 0, SEEK_END) == 0) {
    const long size = std::ftell(f);
    std::rewind(f);
    if (size > 0) {
      s.resize((size_t)size);
      const size_t got = std::fread(&s[0], 1, s.size(), f);
      s.resize(got);
    }
  }
  char buf[1 << 16];  // not seekable (a pipe), or grown since
  size_t got;
  while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, got);
  std::fclose(f);
  return s;
}

}  // namespace fastparse
