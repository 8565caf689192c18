// hammlet_b200 host side — tiny driver for tests/test_host_cli.py: parses a text file of numbers either with
// the input pipeline (FastParse.hpp) or with the reference's extraction loop `while (input >> v)`
// (wavelet.hpp:131) and writes the values as raw float32, so the two can be compared bit for bit.
//   parse_tool fast|slow|auto|text|gz|f32 THREADS IN OUT      prints "<count> <seconds>"
#include <chrono>
#include <cstdio>
#include <iostream>

#include "FastParse.hpp"

int main(int argc, const char* argv[]) {
  try {
    if (argc != 5) throw std::runtime_error("usage: parse_tool fast|slow|auto|text|gz|f32 THREADS IN OUT");
    const std::string mode = argv[1];
    const unsigned threads = (unsigned)std::stoul(argv[2]);
    std::ifstream in(argv[3], std::ios::binary);
    if (!in) throw std::runtime_error("cannot read input");
    std::vector<float> values;
    const auto t0 = std::chrono::steady_clock::now();
    if (mode == "slow") {
      float v;
      while (in >> v) values.push_back(v);
    } else if (mode == "fast") {
      const std::string text = fastparse::slurp(in);
      fastparse::parseFloats(text.data(), text.size(), values, threads);
    } else {  // auto | text | gz | f32: the formats of the command line's -F
      const std::string bytes = fastparse::slurpFile(argv[3]);
      fastparse::parseAny(bytes.data(), bytes.size(), fastparse::formatFromName(mode), values, threads);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    FILE* f = std::fopen(argv[4], "wb");
    if (!f) throw std::runtime_error("cannot write output");
    if (!values.empty()) std::fwrite(values.data(), sizeof(float), values.size(), f);
    std::fclose(f);
    std::cout << values.size() << " " << secs << std::endl;
    return 0;
  } catch (std::exception& e) {
    std::cerr << "[ERROR] " << e.what() << std::endl;
    return 1;
  }
}
