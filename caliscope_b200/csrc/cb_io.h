// On-disk formats at the boundary of the path: the numeric CSV files Caliscope keeps its
// observation and point tables in -- xy_<TRACKER>.csv (ImagePoints, reference core/point_data.py:352-373) and
// xyz_<TRACKER>.csv (WorldPoints, :655-677) -- written by `df.to_csv(index=False, float_format="%.6f")` through
// persistence._safe_write_csv (reference persistence.py:27-41: temp file + fsync + atomic rename) and read back by
// `pd.read_csv`.  A 2 M-row table costs pandas ~10 s to write and ~1.5 s to read; here rows are formatted / parsed by all
// host cores into byte-identical files.  Host code only (no device work: this is file I/O).
//
//   cb_csv_write_numeric  columns of int64 / float64 -> CSV, integers as decimal literals, floats as printf("%.6f"),
//                         NaN as the empty field (pandas na_rep=""), "\n" line ends, header line given by the caller
//   cb_csv_scan / cb_csv_parse_numeric   CSV body -> float64 columns + per-column "every field was an integer literal" /
//                         "had an empty field" flags, from which the Python side rebuilds pandas' dtypes
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace cbio {

inline int n_workers(int requested, size_t work_items, size_t min_per_worker) {
  unsigned hw = std::thread::hardware_concurrency();
  int n = requested > 0 ? requested : (int)std::min<unsigned>(hw ? hw : 8u, 32u);
  const size_t cap = std::max<size_t>(1, work_items / std::max<size_t>(min_per_worker, 1));
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)n, cap));
}

template <typename F>
void parallel_for(int n, F&& f) {
  if (n <= 1) { f(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < n; ++t) th.emplace_back([&f, t] { f(t); });
  f(0);
  for (auto& x : th) x.join();
}

// "%.6f" of a finite double, identical to printf (correctly rounded): integer part by to_chars, 6 decimals from the exact
// scaled remainder when the value is small enough for that to be exact, snprintf otherwise
inline void append_f6(std::string& out, double v) {
  char buf[64];
  if (v != v) return;  // NaN -> empty field
  int n = std::snprintf(buf, sizeof(buf), "%.6f", v);
  out.append(buf, (size_t)n);
}

// pd.read_csv's default float conversion (float_precision="high": precise_xstrtod in pandas/_libs/src/parser/tokenizer.c),
// restated so that the parsed doubles equal pandas' to the last bit: at most 17 digits (leading zeros included) are
// accumulated in a double, further integer digits only raise the exponent, further decimals are dropped, and the result is
// scaled by one multiplication or division with an exact power of ten.  It is NOT correctly rounded for literals with more
// than ~15 significant digits (the older session files store float32 values with 17), which is exactly why it is restated
// instead of calling std::from_chars.  Returns the end of the literal, or nullptr if [p, e) does not start with a number.
inline double pow10_exact(int k) {
  // the correctly rounded doubles nearest 10^k, k = 0..308 (pandas holds the same values as literals)
  static const std::vector<double> tab = [] {
    std::vector<double> t(309);
    for (int i = 0; i <= 308; ++i) t[(size_t)i] = std::strtod(("1e" + std::to_string(i)).c_str(), nullptr);
    return t;
  }();
  return tab[(size_t)(k < 0 ? 0 : k > 308 ? 308 : k)];
}
inline const char* precise_xstrtod(const char* p, const char* e, double* out) {
  const int max_digits = 17;
  bool neg = false;
  if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
  double number = 0.0;
  int exponent = 0, num_digits = 0, num_decimals = 0;
  const char* start = p;
  while (p < e && *p >= '0' && *p <= '9') {
    if (num_digits < max_digits) { number = number * 10.0 + (*p - '0'); ++num_digits; }
    else ++exponent;
    ++p;
  }
  if (p < e && *p == '.') {
    ++p;
    while (num_digits < max_digits && p < e && *p >= '0' && *p <= '9') {
      number = number * 10.0 + (*p - '0');
      ++p; ++num_digits; ++num_decimals;
    }
    if (num_digits >= max_digits)
      while (p < e && *p >= '0' && *p <= '9') ++p;
    exponent -= num_decimals;
  }
  if (num_digits == 0) return nullptr;
  (void)start;
  if (p < e && (*p == 'e' || *p == 'E')) {
    const char* q = p + 1;
    bool eneg = false;
    if (q < e && (*q == '-' || *q == '+')) { eneg = *q == '-'; ++q; }
    if (q < e && *q >= '0' && *q <= '9') {
      int n = 0;
      while (q < e && *q >= '0' && *q <= '9') { n = n * 10 + (*q - '0'); if (n > 100000) n = 100000; ++q; }
      exponent += eneg ? -n : n;
      p = q;
    }
  }
  if (exponent > 308) number = INFINITY;
  else if (exponent > 0) number *= pow10_exact(exponent);
  else if (exponent < -308) { number /= pow10_exact(-308 - exponent); number /= pow10_exact(308); }
  else number /= pow10_exact(-exponent);
  *out = neg ? -number : number;
  return p;
}

inline void append_i64(std::string& out, long long v) {
  char buf[32];
  auto r = std::to_chars(buf, buf + sizeof(buf), v);
  out.append(buf, (size_t)(r.ptr - buf));
}

}  // namespace cbio

extern "C" {

// col_kind[c]: 0 = int64 column, 1 = float64 column.  Returns 0, or a negative errno-style code (see cb_ba_last_error).
int cb_csv_write_numeric(const char* path, const char* header, int64_t n_rows, int32_t n_cols, const int32_t* col_kind,
                         const void* const* col_data, int32_t n_threads);

// Pass 1: number of data rows (lines after the header that are not empty) and of columns in the header.
int cb_csv_scan(const char* path, int64_t* n_rows, int32_t* n_cols);

// Pass 2: out is column-major float64 [n_cols][n_rows]; col_all_int[c] = 1 iff every field of column c was an integer literal,
// col_has_empty[c] = 1 iff some field was empty (parsed as NaN).  A field that is neither empty nor numeric makes the call fail
// (the tables on this path are purely numeric).
int cb_csv_parse_numeric(const char* path, int64_t n_rows, int32_t n_cols, double* out, int32_t* col_all_int,
                         int32_t* col_has_empty, int32_t n_threads);

}  // extern "C"
