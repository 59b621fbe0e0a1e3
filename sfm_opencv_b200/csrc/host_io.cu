// host_io.cu -- host-only C++: the output files of the reference, byte for byte.
//
//   sfm_save_structure    == save_structure()     OpenCV_SFM/NViewReconstuct.cpp:186-227
//                            (cv::FileStorage YAML 1.0: "Camera Count", "Point Count",
//                            "Rotations", "Motions", "Points", "Colors")
//   sfm_write_ply_binary  == write_ply_binary()   OpenCV_SFM/NViewReconstuct.cpp:229-294
//                            (27 bytes per vertex, NaN vertices skipped)
//
// cv::FileStorage is not available to a library without OpenCV, so the part of its YAML
// emitter that save_structure() exercises is restated here (OpenCV 4.x
// modules/core/src/persistence.cpp / persistence_yml.cpp; the reference pins 4.4.0): block
// and flow collections, "!!opencv-matrix" maps, the 71-column wrap margin, the
// "%d." / "%.16e" number format.  tests/test_host_io.py regenerates the reference's bundled
// Viewer/structure.yml, structure_ba.yml and structure_ba.ply from their parsed contents and
// compares the bytes.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sfm_b200.h"

namespace {

// ---- minimal restatement of cv::FileStorage's YAML emitter --------------------------------
class YamlEmitter {
 public:
  explicit YamlEmitter(FILE* f) : f_(f) {
    fputs("%YAML:1.0\n---\n", f_);
    stack_.push_back({0, kMap | kEmpty});
  }
  // key may be null inside sequences
  void scalar(const char* key, const char* data) {
    Struct& cur = stack_.back();
    const size_t keylen = key ? strlen(key) : 0, datalen = data ? strlen(data) : 0;
    if (cur.flags & kFlow) {
      if (!(cur.flags & kEmpty)) buf_ += ',';
      const int new_offset = static_cast<int>(buf_.size() + keylen + datalen);
      if (new_offset > kWrapMargin && static_cast<int>(buf_.size()) > cur.indent)
        flush();
      else
        buf_ += ' ';
    } else {
      flush();
      if (!(cur.flags & kMap)) {
        buf_ += '-';
        if (data) buf_ += ' ';
      }
    }
    if (key) {
      buf_.append(key, keylen);
      buf_ += ':';
      if (!(cur.flags & kFlow) && data) buf_ += ' ';
    }
    if (data) buf_.append(data, datalen);
    stack_.back().flags &= ~kEmpty;
  }
  void begin(const char* key, bool map, bool flow, const char* type_name) {
    char tmp[64];
    const char* data = nullptr;
    if (flow) {
      const char c = map ? '{' : '[';
      if (type_name) snprintf(tmp, sizeof tmp, "!!%s %c", type_name, c);
      else { tmp[0] = c; tmp[1] = 0; }
      data = tmp;
    } else if (type_name) {
      snprintf(tmp, sizeof tmp, "!!%s", type_name);
      data = tmp;
    }
    scalar(key, data);
    Struct s;
    s.indent = stack_.back().indent;
    s.flags = (map ? kMap : 0) | (flow ? kFlow : 0) | kEmpty;
    if (!(stack_.back().flags & kFlow)) s.indent += 3 + (flow ? 1 : 0);
    stack_.push_back(s);
  }
  void end() {
    const Struct cur = stack_.back();
    if (cur.flags & kFlow) {
      if (static_cast<int>(buf_.size()) > cur.indent && !(cur.flags & kEmpty)) buf_ += ' ';
      buf_ += (cur.flags & kMap) ? '}' : ']';
    } else if (cur.flags & kEmpty) {
      flush();                                  // own line, at the collection's indentation
      buf_ += (cur.flags & kMap) ? "{}" : "[]";
    }
    stack_.pop_back();
  }
  void integer(const char* key, int v) {
    char b[32];
    snprintf(b, sizeof b, "%d", v);
    scalar(key, b);
  }
  void real(const char* key, double v) {
    char b[64];
    format_double(b, sizeof b, v);
    scalar(key, b);
  }
  // cv::write(fs, name, Mat) for a CV_64F matrix
  void matrix_f64(const char* key, int rows, int cols, const double* data) {
    begin(key, true, false, "opencv-matrix");
    integer("rows", rows);
    integer("cols", cols);
    scalar("dt", "d");
    begin("data", false, true, nullptr);
    for (int i = 0; i < rows * cols; ++i) real(nullptr, data[i]);
    end();
    end();
  }
  void finish() {
    while (stack_.size() > 1) end();
    flush();
  }

 private:
  enum { kMap = 1, kFlow = 2, kEmpty = 4 };
  static constexpr int kWrapMargin = 71;
  struct Struct { int indent; int flags; };

  // emits the pending line (if it holds more than indentation) and starts a new one at the
  // indentation of the current collection
  void flush() {
    if (static_cast<int>(buf_.size()) > space_) {
      buf_ += '\n';
      fputs(buf_.c_str(), f_);
    }
    space_ = stack_.back().indent;
    buf_.assign(static_cast<size_t>(space_), ' ');
  }
  // printf("%.16e") as the reference's C runtime (MSVC) prints it.  glibc differs in exactly
  // one case: when the decimal expansion of the double ends in an exact ...5 at the 18th
  // significant digit (typical for float32 values widened to double, i.e. every coordinate
  // of "Points" before bundle adjustment) glibc rounds half to even, the reference's runtime
  // half away from zero (bundled Viewer/structure.yml: -1.85865020751953125 is printed as
  // -1.8586502075195313e+00).  Ties are detected on the exact expansion.
  static void format_e16(char* buf, size_t n, double value) {
    snprintf(buf, n, "%.16e", value);
    char exact[1200];
    snprintf(exact, sizeof exact, "%.1100e", value);      // exact: a double has < 1100 digits
    const char* p = exact + (exact[0] == '-');            // d.ddddd...e+XX
    // digits after the point: p[2 + k]; the 17th significant digit is p[2 + 15], the 18th p[2 + 16]
    if (p[2 + 16] != '5') return;
    for (const char* q = p + 2 + 17; *q && *q != 'e'; ++q)
      if (*q != '0') return;
    if ((p[2 + 15] - '0') & 1) return;                    // glibc already rounded up (to even)
    // tie that glibc rounded down: add one unit in the last printed place
    char* m = buf + (buf[0] == '-');
    int i = 2 + 15;
    for (;;) {
      if (m[i] == '.') { --i; continue; }
      if (m[i] != '9') { ++m[i]; break; }
      m[i] = '0';
      if (i == 0) break;                                   // cannot happen: last digit is even
      --i;
    }
  }
  // cv::fs::doubleToString
  static void format_double(char* buf, size_t n, double value) {
    uint64_t u;
    memcpy(&u, &value, 8);
    const unsigned hi = static_cast<unsigned>(u >> 32);
    if ((hi & 0x7ff00000u) != 0x7ff00000u) {
      const int iv = static_cast<int>(lrint(value));       // cvRound
      if (static_cast<double>(iv) == value) snprintf(buf, n, "%d.", iv);
      else format_e16(buf, n, value);
    } else {
      const unsigned lo = static_cast<unsigned>(u);
      if ((hi & 0x7fffffffu) + (lo != 0) > 0x7ff00000u) snprintf(buf, n, ".Nan");
      else snprintf(buf, n, static_cast<int64_t>(u) < 0 ? "-.Inf" : ".Inf");
    }
  }

  FILE* f_;
  std::string buf_;
  int space_ = 0;
  std::vector<Struct> stack_;
};

}  // namespace

extern "C" {

int sfm_save_structure(const char* file_name, int n_cam, const double* rotations,
                       const double* motions, int64_t n_pts, const double* structure,
                       int64_t n_colors, const uint8_t* colors) {
  if (!file_name || n_cam < 0 || n_pts < 0 || n_colors < 0 || (n_cam > 0 && (!rotations || !motions)) ||
      (n_pts > 0 && !structure) || (n_colors > 0 && !colors))
    return SFM_E_INVALID;
  FILE* f = fopen(file_name, "wb");
  if (!f) return SFM_E_INVALID;
  YamlEmitter y(f);
  y.integer("Camera Count", n_cam);
  y.integer("Point Count", static_cast<int>(n_pts));
  y.begin("Rotations", false, false, nullptr);
  for (int i = 0; i < n_cam; ++i) y.matrix_f64(nullptr, 3, 3, rotations + 9 * i);
  y.end();
  y.begin("Motions", false, false, nullptr);
  for (int i = 0; i < n_cam; ++i) y.matrix_f64(nullptr, 3, 1, motions + 3 * i);
  y.end();
  y.begin("Points", false, false, nullptr);
  for (int64_t i = 0; i < n_pts; ++i) {          // Point3d: flow sequence of 3 doubles
    y.begin(nullptr, false, true, nullptr);
    for (int k = 0; k < 3; ++k) y.real(nullptr, structure[3 * i + k]);
    y.end();
  }
  y.end();
  y.begin("Colors", false, false, nullptr);
  for (int64_t i = 0; i < n_colors; ++i) {       // Vec3b: flow sequence of 3 ints
    y.begin(nullptr, false, true, nullptr);
    for (int k = 0; k < 3; ++k) y.integer(nullptr, colors[3 * i + k]);
    y.end();
  }
  y.end();
  y.finish();
  const bool ok = ferror(f) == 0;
  return (fclose(f) == 0 && ok) ? SFM_OK : SFM_E_INVALID;
}

int sfm_write_ply_binary(const char* path, int64_t n, const float* xyz_normal, const uint8_t* rgb,
                         int crlf) {
  if (!path || n < 0 || (n > 0 && (!xyz_normal || !rgb))) return SFM_E_INVALID;
  int64_t valid = 0;
  auto bad = [&](int64_t i) {
    for (int k = 0; k < 6; ++k)
      if (isnan(xyz_normal[6 * i + k])) return true;
    return false;
  };
  for (int64_t i = 0; i < n; ++i) valid += !bad(i);
  FILE* f = fopen(path, "wb");
  if (!f) return SFM_E_INVALID;
  const char* nl = crlf ? "\r\n" : "\n";          // the reference's text-mode header on Windows
  fprintf(f, "ply%sformat binary_little_endian 1.0%selement vertex %lld%s", nl, nl,
          static_cast<long long>(valid), nl);
  const char* props[] = {"float x", "float y", "float z", "float nx", "float ny", "float nz",
                         "uchar red", "uchar green", "uchar blue"};
  for (const char* p : props) fprintf(f, "property %s%s", p, nl);
  fprintf(f, "end_header%s", nl);
  for (int64_t i = 0; i < n; ++i) {
    if (bad(i)) continue;
    fwrite(xyz_normal + 6 * i, sizeof(float), 6, f);
    fwrite(rgb + 3 * i, 1, 3, f);
  }
  const bool ok = ferror(f) == 0;
  return (fclose(f) == 0 && ok) ? SFM_OK : SFM_E_INVALID;
}

}  // extern "C"
