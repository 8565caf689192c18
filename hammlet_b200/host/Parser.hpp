// hammlet_b200 host side — command-line parser with the reference's token-group semantics
// (reference: src/Parser.hpp).  The command line is a flat token list; a registered flag opens a
// group and every following non-flag token belongs to it.  No positional arguments, no "-abc"
// stacking, a flag given twice is an error, defaults are replaced as soon as the flag appears.
#pragma once

#include <initializer_list>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

template <typename T>
T convertType(const std::string& s) {
  T x;
  if (std::istringstream(s) >> x) return x;
  throw std::runtime_error("Conversion failed for string \"" + s + "\"!");
}

inline std::vector<std::string> tokenize(const std::string& source, const char* delimiter = " ", bool keepEmpty = false) {
  std::vector<std::string> out;
  size_t from = 0;
  while (from <= source.size()) {
    size_t to = source.find_first_of(delimiter, from);
    if (to == std::string::npos) to = source.size();
    if (keepEmpty || to > from) out.push_back(source.substr(from, to - from));
    from = to + 1;
  }
  return out;
}

class Parser {
  struct Group {
    std::vector<std::string> names;
    std::vector<std::string> tokens;
    bool given = false;
  };
  std::vector<std::string> mInput;
  std::vector<Group> mGroups;
  std::map<std::string, size_t> mIndex;
  std::set<std::string> mBlocked;
  bool mParsed = false;

  const Group& group(const std::string& flag) const {
    if (!mParsed) throw std::runtime_error("Command line has not been parsed yet!");
    auto it = mIndex.find(flag);
    if (it == mIndex.end()) throw std::runtime_error(flag + " is not registered as a flag!");
    return mGroups[it->second];
  }

 public:
  Parser(int argc, const char* argv[]) : mInput(argv + 1, argv + argc) {}
  explicit Parser(std::vector<std::string> tokens) : mInput(std::move(tokens)) {}

  void registerFlags(std::initializer_list<std::string> flags, const std::string& defaults = "") {
    if (mParsed) throw std::runtime_error("Cannot register flags, tokens have already been parsed!");
    Group g;
    for (const std::string& f : flags) {
      if (mIndex.count(f)) throw std::runtime_error("Flag " + f + " has already been registered!");
      if (mBlocked.count(f)) throw std::runtime_error("Flag " + f + " is blocked!");
      mIndex[f] = mGroups.size();
      g.names.push_back(f);
    }
    g.tokens = tokenize(defaults);
    mGroups.push_back(g);
  }

  void parseArgs() {
    mParsed = true;
    if (mInput.empty()) return;
    if (!mIndex.count(mInput[0]))
      throw std::runtime_error("First input token (" + mInput[0] +
                               ") is not a registered flag; parser does not support positional arguments!");
    Group* open = nullptr;
    for (const std::string& tok : mInput) {
      auto it = mIndex.find(tok);
      if (it != mIndex.end()) {
        open = &mGroups[it->second];
        if (open->given) throw std::runtime_error("Duplicate flag " + tok + "!");
        open->given = true;
        open->tokens.clear();
      } else {
        open->tokens.push_back(tok);
      }
    }
    mInput.clear();
  }

  template <class T>
  T parse(const std::string& flag, size_t index = 0) const {
    const Group& g = group(flag);
    if (index >= g.tokens.size()) throw std::runtime_error("Not enough arguments for flag " + flag + "!");
    return convertType<T>(g.tokens[index]);
  }

  template <class T>
  std::vector<T> parseVector(const std::string& flag, size_t begin = 0, size_t end = 0) const {
    const Group& g = group(flag);
    if (end == 0) end = g.tokens.size();
    if (end <= begin) throw std::runtime_error("Invalid range for flag " + flag + "!");
    if (end > g.tokens.size()) throw std::runtime_error("Not enough arguments for flag " + flag + "!");
    std::vector<T> out;
    for (size_t i = begin; i < end; ++i) out.push_back(convertType<T>(g.tokens[i]));
    return out;
  }

  bool isSet(const std::string& flag) const {
    auto it = mIndex.find(flag);
    if (it == mIndex.end()) throw std::runtime_error(flag + " is not registered as a flag!");
    return mGroups[it->second].given;
  }

  size_t nrTokens(const std::string& flag) const { return group(flag).tokens.size(); }
  std::vector<std::string> tokens(const std::string& flag) const { return group(flag).tokens; }

  void blockFlag(const std::string& flag) { mBlocked.insert(flag); }

  // parser over the tokens of one flag; all of this parser's flags are blocked in it
  Parser subparser(const std::string& flag) const {
    Parser sub(tokens(flag));
    for (const auto& kv : mIndex) sub.blockFlag(kv.first);
    return sub;
  }

  // -g: one line per flag group: "[*] -f -input-file : tokens"  ([ ] if not given)
  void print() const {
    if (!mParsed) throw std::runtime_error("Command line has not been parsed yet!");
    for (size_t i = 0; i < mGroups.size(); ++i) {
      std::vector<std::string> names;
      for (const auto& kv : mIndex)
        if (kv.second == i) names.push_back(kv.first);  // map order, like the reference
      std::cout << (mGroups[i].given ? "[*]" : "[ ]");
      for (const auto& n : names) std::cout << " " << n;
      std::cout << " :";
      for (const auto& t : mGroups[i].tokens) std::cout << " " << t;
      std::cout << std::endl;
    }
  }
};
