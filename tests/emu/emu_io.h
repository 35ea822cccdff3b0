// Raw binary file helpers shared by the emulator harnesses.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

template <typename T>
static std::vector<T> ReadAll(const std::string& path, bool optional = false) {
  std::vector<T> out;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) {
    if (optional) return out;
    std::fprintf(stderr, "cannot open %s\n", path.c_str());
    std::exit(2);
  }
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize((size_t)n / sizeof(T));
  if (n && std::fread(out.data(), 1, (size_t)n, f) != (size_t)n) std::exit(2);
  std::fclose(f);
  return out;
}
template <typename T>
static void WriteAll(const std::string& path, const std::vector<T>& v) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) std::exit(2);
  std::fwrite(v.data(), sizeof(T), v.size(), f);
  std::fclose(f);
}
