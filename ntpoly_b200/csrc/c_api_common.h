// Shared by the extern "C" translation units: opaque handles and the host-side MatrixMarket reader.
#pragma once
#include "../../include/ntpoly_b200.h"
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace ntb {
namespace capi {
// opaque handle = caller-owned int[12]; first 8 bytes carry the object pointer
// (the reference TRANSFERs a derived type holding one POINTER, WrapperModule.F90:8)
template <typename T> T* get(const int* ih) {
  T* p;
  std::memcpy(&p, ih, sizeof(p));
  NTB_CHECK(p != nullptr, "null handle passed to ntpoly_b200");
  return p;
}
template <typename T> void put(int* ih, T* p) {
  std::memset(ih, 0, NTB_SIZE_wrp * sizeof(int));
  std::memcpy(ih, &p, sizeof(p));
}
inline void clear(int* ih) { std::memset(ih, 0, NTB_SIZE_wrp * sizeof(int)); }

// ---- MatrixMarket (host side; PSMatrixModule.F90:351-570, sparse_includes/ConstructMatrixFromFile.f90)
struct MMData {
  int n = 0;          // rows
  int ncols = 0;
  bool is_complex = false;
  std::vector<int> rows, cols;
  std::vector<double> re, im;
};
// every rank parses the file, rank `rank` of `size` keeps a disjoint share of the entries
inline MMData read_matrix_market(const std::string& path, int rank, int size) {
  std::ifstream f(path);
  NTB_CHECK(f.good(), "cannot open MatrixMarket file");
  std::string line;
  std::getline(f, line);
  std::string lower = line;
  std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
  MMData d;
  d.is_complex = lower.find("complex") != std::string::npos;
  const bool pattern = lower.find("pattern") != std::string::npos;
  const bool symmetric = lower.find(" symmetric") != std::string::npos;
  const bool skew = lower.find("skew-symmetric") != std::string::npos;
  const bool hermitian = lower.find("hermitian") != std::string::npos;
  while (std::getline(f, line)) if (!line.empty() && line[0] != '%') break;
  long long nr = 0, nc = 0, nnz = 0;
  { std::istringstream ss(line); ss >> nr >> nc >> nnz; }
  d.n = (int)nr;
  d.ncols = (int)nc;
  for (long long i = 0; i < nnz; ++i) {
    int r, c;
    double vr = 1.0, vi = 0.0;
    f >> r >> c;
    if (!pattern) { f >> vr; if (d.is_complex) f >> vi; }
    if ((i % size) != rank) continue;
    d.rows.push_back(r); d.cols.push_back(c); d.re.push_back(vr); d.im.push_back(vi);
    if ((symmetric || skew || hermitian) && r != c) {
      d.rows.push_back(c); d.cols.push_back(r);
      d.re.push_back(skew ? -vr : vr);
      d.im.push_back(hermitian ? -vi : (skew ? -vi : vi));
    }
  }
  return d;
}
}  // namespace capi
}  // namespace ntb
