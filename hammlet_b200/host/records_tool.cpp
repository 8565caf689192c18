// hammlet_b200 host side — tiny driver that replays recorded iterations through Records, used by
// tests/test_host_records.py to check the record files byte for byte against the reference's.
// Input (stdin): T K nsweeps, then per sweep: B, B block sizes, B states.  Output: files PREFIX*SUFFIX.
#include "Records.hpp"

int main(int argc, const char* argv[]) {
  try {
    if (argc != 3) throw std::runtime_error("usage: records_tool PREFIX SUFFIX < iterations");
    size_t T, K, n;
    std::cin >> T >> K >> n;
    Records rec(T, argv[1], argv[2], K);
    rec.setRecordStateSequence(true, true);
    rec.setRecordBlocks(true, true);
    rec.setRecordCompression(true, true);
    rec.setRecordMarginals(true, true);
    rec.setRecordSegments(true, true);
    for (size_t it = 0; it < n; ++it) {
      size_t B;
      std::cin >> B;
      std::vector<size_t> sizes(B), states(B);
      for (auto& v : sizes) std::cin >> v;
      for (auto& v : states) std::cin >> v;
      for (size_t b = 0; b < B; ++b) rec.record(states[b], sizes[b]);
    }
    rec.close();
    return 0;
  } catch (std::exception& e) {
    std::cerr << "[ERROR] " << e.what() << std::endl;
    return 1;
  }
}
