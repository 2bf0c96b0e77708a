// houdini_file_saver.cpp -- ASCII Houdini .geo ("PGEOMETRY V5") particle frames, off the critical path.
//
// Produces the same bytes as the reference's libclsph/file_save_delegates/houdini_file_saver.cpp
// :25-92 through util/houdini_geo/HoudiniFileDumpHelper.cpp:19-90: header, per point
// "x y z 0 (vx vy vz<TAB>r g b<TAB>mass)", the single particle primitive, trailer. Numbers use
// the default iostream float format (= printf %g, 6 significant digits).
//
// Once the step runs on the GPU, frame export is by far the largest wall-time term (SURVEY 8f
// row 2): ~100 bytes of text per particle per frame against ~9 ms of simulation per frame at 1 Mi
// particles. So writeFrameToFile only copies the seven floats per particle it needs (28 of the 80
// bytes) into a job and returns; a background thread formats the job on all host cores (put_g
// below: exact integer formatting, ~5x faster than snprintf) and writes the file. At most two frames are in flight; wait() or the
// destructor blocks until everything is on disk. `asynchronous = false` restores the reference's
// "written when the call returns".
//
// The reference's other output, chosen there at compile time with USE_PARTIO (:78-88 through
// util/partio/PartioFunctions.h:5-65): a binary Houdini "Bgeo V5" file written by libpartio's writer
// with the attributes position, velocity, color, id, mass, pscale. libpartio is not vendored by the
// reference (headers only), so write_job_bgeo below restates the layout of Partio's published BGEO
// writer (partio 1.1, src/lib/io/BGEO.cpp: everything big-endian; header of the magic, 'V', 5 and
// nine counts; one definition per attribute other than position; per point x y z 1 and the
// attribute words; the "generator"/"papi" primitive attribute; one particle-system primitive
// 0x8000 listing every point; 0x00 0xff). PARITY UNPINNED: there is no libpartio here to compare
// bytes with; tests/test_bgeo.py reads the file back with an independent parser of that layout.
// One deliberate difference: the reference fills velocity[0] three times and
// leaves [1] and [2] as libpartio's allocator left them; all three components are written here.
#include "file_save_delegates/houdini_file_saver.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

// Frame number as 7 digits: 9-wide zero padding cut to its last 7 characters (reference :15-23).
std::string frame_suffix(int num) {
  std::ostringstream ss;
  ss << std::setw(9) << std::setfill('0') << num;
  std::string s = ss.str();
  return s.size() > 7 ? s.substr(s.size() - 7) : s;
}

// Density -> colour ramp of the reference (:47-60).
void density_colour(float rho, float* r, float* g, float* b) {
  *r = (rho > 1000.f && rho <= 2000.f) ? (rho - 1000.f) / 1000.f : 0.f;
  *g = (rho >= 0.f && rho < 1000.f) ? 1.f - rho / 1000.f : 0.f;
  *b = (rho >= 500.f && rho <= 1000.f) ? (rho - 500.f) / 500.f
       : (rho >= 1000.f && rho <= 1500.f) ? 1.f - (rho - 1000.f) / 500.f
                                          : 0.f;
}

// ---- printf("%g", (double)v) for a float, byte for byte, with integer arithmetic --------------------
// A frame is ~11 numbers per particle; number formatting is the whole cost of the writer. A float is
// m * 2^q exactly (m < 2^24), so its six significant decimal digits, rounded half-to-even on the exact
// value like glibc's printf, come out of one 128-bit multiply by a power of ten and a shift. That covers
// 1e-24 <= |v| < 1e6 (everything a simulation produces); the rest goes to std::to_chars, which is
// specified to match printf. Checked against snprintf("%g") for ALL 2^32 float bit patterns
// (tests/test_host_api.py runs a strided sample of the same comparison).
static const unsigned __int128 kPow10[30] = {
    (unsigned __int128)1ull, (unsigned __int128)10ull, (unsigned __int128)100ull, (unsigned __int128)1000ull,
    (unsigned __int128)10000ull, (unsigned __int128)100000ull, (unsigned __int128)1000000ull, (unsigned __int128)10000000ull,
    (unsigned __int128)100000000ull, (unsigned __int128)1000000000ull, (unsigned __int128)10000000000ull,
    (unsigned __int128)100000000000ull, (unsigned __int128)1000000000000ull, (unsigned __int128)10000000000000ull,
    (unsigned __int128)100000000000000ull, (unsigned __int128)1000000000000000ull, (unsigned __int128)10000000000000000ull,
    (unsigned __int128)100000000000000000ull, (unsigned __int128)1000000000000000000ull,
    (unsigned __int128)10000000000000000000ull, (unsigned __int128)10000000000000000000ull * 10u,
    (unsigned __int128)10000000000000000000ull * 100u, (unsigned __int128)10000000000000000000ull * 1000u,
    (unsigned __int128)10000000000000000000ull * 10000u, (unsigned __int128)10000000000000000000ull * 100000u,
    (unsigned __int128)10000000000000000000ull * 1000000u, (unsigned __int128)10000000000000000000ull * 10000000u,
    (unsigned __int128)10000000000000000000ull * 100000000u, (unsigned __int128)10000000000000000000ull * 1000000000u,
    (unsigned __int128)10000000000000000000ull * 10000000000ull};

// |v| = m * 2^q exactly. Six significant decimal digits, correctly rounded (half to even on the exact
// value, which is what glibc's printf does), for 1e-24 <= |v| < 1e6; returns false outside that range.
inline bool digits6(uint32_t m, int q, double dv, uint32_t* digits, int* k_out) {  // dv = m * 2^q
  // decimal exponent guess from the binary one: |v| in [2^(q+b-1), 2^(q+b)), b = bit length of m
  const int b = 32 - __builtin_clz(m);
  // floor(log10(2^(q+b-1))) = k or k-1 of the true value (78913 / 2^18 = log10(2) to 6 digits; >> floors)
  int k = ((q + b - 1) * 78913) >> 18;
  {
    // settle it against the neighbouring power of ten (the double nearest to 10^k: no float lies strictly
    // between that and the real power, and the loop below still repairs an off-by-one)
    static const double kTen[36] = {1e-27, 1e-26, 1e-25, 1e-24, 1e-23, 1e-22, 1e-21, 1e-20, 1e-19, 1e-18, 1e-17, 1e-16,
                                    1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9,  1e-8,  1e-7,  1e-6,  1e-5,  1e-4,
                                    1e-3,  1e-2,  1e-1,  1e0,   1e1,   1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8};
    if (k >= -26 && k <= 6) {
      if (dv >= kTen[k + 28]) ++k;
    }
  }
  for (int attempt = 0; attempt < 3; ++attempt) {
    const int s = 5 - k;  // scale so that the integer part has six digits
    if (s < 0 || s > 29) return false;
    uint32_t d;
    if (s <= 12 && q < 0 && q > -64) {  // everything fits 64 bits: m * 10^s < 2^24 * 2^40
      static const uint64_t kPow10_64[13] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                                             100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull};
      const uint64_t n64 = (uint64_t)m * kPow10_64[s];
      const int sh = -q;
      const uint64_t whole = n64 >> sh;
      if (whole >= 1000000u) { ++k; continue; }
      if (whole < 100000u) { --k; continue; }
      const uint64_t rem = n64 & ((1ull << sh) - 1u), half = 1ull << (sh - 1);
      d = (uint32_t)whole;
      if (rem > half || (rem == half && (d & 1u))) ++d;
      if (d == 1000000u) { d = 100000u; ++k; }
      *digits = d;
      *k_out = k;
      return true;
    }
    unsigned __int128 n = (unsigned __int128)m * kPow10[s];
    if (q >= 0) {
      if (q > 20) return false;
      n <<= q;
      if (n >= 10000000u) { ++k; continue; }
      d = (uint32_t)n;
      if (d >= 1000000u) { ++k; continue; }
      if (d < 100000u) { --k; continue; }
    } else {
      const int sh = -q;
      if (sh > 126) return false;
      const unsigned __int128 whole = n >> sh;
      if (whole >= 1000000u) { ++k; continue; }
      if (whole < 100000u) { --k; continue; }
      const unsigned __int128 rem = n & ((((unsigned __int128)1) << sh) - 1u);
      const unsigned __int128 half = ((unsigned __int128)1) << (sh - 1);
      d = (uint32_t)whole;
      if (rem > half || (rem == half && (d & 1u))) ++d;
      if (d == 1000000u) { d = 100000u; ++k; }
    }
    *digits = d;
    *k_out = k;
    return true;
  }
  return false;
}

inline char* put_g(char* p, float v) {
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  const uint32_t frac = bits & 0x7fffffu, ex = (bits >> 23) & 0xffu;
  if (ex == 0xffu) return p + std::snprintf(p, 32, "%g", (double)v);  // inf / nan: the C library's spelling
  if (bits & 0x80000000u) *p++ = '-';
  if (ex == 0 && frac == 0) { *p++ = '0'; return p; }
  const uint32_t m = ex ? (frac | 0x800000u) : frac;
  const int q = (ex ? (int)ex : 1) - 150;
  uint32_t d;
  int k;
  const double a = std::fabs((double)v);
  if (!digits6(m, q, a, &d, &k)) {
    return std::to_chars(p, p + 32, a, std::chars_format::general, 6).ptr;
  }
  // six digits through a two-digit table; the copies below write up to 8 bytes past the text (the
  // caller's buffers have that much slack), the pointer only advances over the valid part
  static const char kPairs[201] =
      "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
      "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
  const uint32_t hi = d / 10000u, rest = d - hi * 10000u, mid = rest / 100u, lo = rest - mid * 100u;
  char dig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::memcpy(dig, kPairs + 2 * hi, 2);
  std::memcpy(dig + 2, kPairs + 2 * mid, 2);
  std::memcpy(dig + 4, kPairs + 2 * lo, 2);
  int last;  // index of the last non-zero digit: %g strips trailing zeros
  if (lo) last = (dig[5] != '0') ? 5 : 4;
  else if (mid) last = (dig[3] != '0') ? 3 : 2;
  else last = (dig[1] != '0') ? 1 : 0;
  if (k >= 0 && k < 6) {
    std::memcpy(p, dig, 8);
    p += k + 1;
    if (last > k) {
      *p++ = '.';
      std::memcpy(p, dig + k + 1, 8 - (k + 1) > 5 ? 5 : 8 - (k + 1));
      p += last - k;
    }
  } else if (k < 0 && k >= -4) {
    std::memcpy(p, "0.0000", 6);
    p += 1 - k;
    std::memcpy(p, dig, 8);
    p += last + 1;
  } else {
    *p++ = dig[0];
    if (last > 0) {
      *p++ = '.';
      std::memcpy(p, dig + 1, 5);
      p += last;
    }
    *p++ = 'e';
    int e = k;
    if (e < 0) { *p++ = '-'; e = -e; } else *p++ = '+';
    if (e >= 100) { *p++ = (char)('0' + e / 100); e %= 100; }
    std::memcpy(p, kPairs + 2 * e, 2);
    p += 2;
  }
  return p;
}

inline char* put_uint(char* p, unsigned int v) { return std::to_chars(p, p + 16, v).ptr; }

struct FramePoint {  // what a frame needs of a particle: 28 of its 80 bytes
  float px, py, pz, vx, vy, vz, rho;
};

struct FrameJob {
  std::string file_name;
  float mass = 0.f;
  float pscale = 0.f;   // support radius h (.bgeo only)
  bool bgeo = false;
  std::vector<FramePoint> points;
};

// A chunk of output text in an uninitialised buffer (std::string::resize would zero-fill ~150 MB per frame).
struct Text {
  std::unique_ptr<char[]> data;
  size_t size = 0;
};

// Text of points [i0, i1): one line each.
void format_points(const FrameJob& job, size_t i0, size_t i1, Text* out) {
  out->data.reset(new char[(i1 - i0) * 160 + 64]);  // 11 numbers of at most 12 characters + 14 separators
  char* p = out->data.get();
  for (size_t i = i0; i < i1; ++i) {
    const FramePoint& q = job.points[i];
    float r, g, b;
    density_colour(q.rho, &r, &g, &b);
    p = put_g(p, q.px); *p++ = ' ';
    p = put_g(p, q.py); *p++ = ' ';
    p = put_g(p, q.pz); *p++ = ' ';
    *p++ = '0'; *p++ = ' '; *p++ = '(';
    p = put_g(p, q.vx); *p++ = ' ';
    p = put_g(p, q.vy); *p++ = ' ';
    p = put_g(p, q.vz); *p++ = '\t';
    p = put_g(p, r); *p++ = ' ';
    p = put_g(p, g); *p++ = ' ';
    p = put_g(p, b); *p++ = '\t';
    p = put_g(p, job.mass); *p++ = ')'; *p++ = '\n';
  }
  out->size = static_cast<size_t>(p - out->data.get());
}

// " i0 i0+1 ... i1-1": the vertex list of the single particle primitive.
void format_indices(size_t i0, size_t i1, Text* out) {
  out->data.reset(new char[(i1 - i0) * 11 + 64]);  // a space and at most 10 digits
  char* p = out->data.get();
  for (size_t i = i0; i < i1; ++i) {
    *p++ = ' ';
    p = put_uint(p, static_cast<unsigned int>(i));
  }
  out->size = static_cast<size_t>(p - out->data.get());
}

void write_job(const FrameJob& job, int threads) {
  const size_t n = job.points.size();
  const size_t parts = std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(std::max(1, threads)), (n + 4095) / 4096));
  std::vector<Text> point_text(parts), index_text(parts);
  auto work = [&](size_t k) {
    const size_t i0 = n * k / parts, i1 = n * (k + 1) / parts;
    format_points(job, i0, i1, &point_text[k]);
    format_indices(i0, i1, &index_text[k]);
  };
  std::vector<std::thread> pool;
  for (size_t k = 1; k < parts; ++k) pool.emplace_back(work, k);
  work(0);
  for (std::thread& t : pool) t.join();

  char line[64];
  std::string head = "PGEOMETRY V5\n";
  std::snprintf(line, sizeof(line), "NPoints %d NPrims 1\n", static_cast<int>(n));
  head += line;
  head += "NPointGroups 0 NPrimGroups 1\n";
  head += "NPointAttrib 3 NVertexAttrib 0 NPrimAttrib 2 NAttrib 0\n";
  head += "PointAttrib\nv 3 float 1 1 1\ncolor 3 float 1 1 1\nmass 1 float 1\n";
  std::string mid = "PrimitiveAttrib\ngenerator 1 index 1 location1\ndopobject 1 index 1 /obj/AutoDopNetwork:1\n";
  std::snprintf(line, sizeof(line), "Part %d", static_cast<int>(n));
  mid += line;
  const std::string tail = " [0\t0]\nbox_object1 unordered\n1 1\nbeginExtra\nendExtra\n";

  std::FILE* f = std::fopen(job.file_name.c_str(), "wb");
  if (!f) {
    std::cerr << "Error while writing to " << job.file_name << std::endl;
    return;
  }
  std::fwrite(head.data(), 1, head.size(), f);
  for (const Text& s : point_text) std::fwrite(s.data.get(), 1, s.size, f);
  std::fwrite(mid.data(), 1, mid.size(), f);
  for (const Text& s : index_text) std::fwrite(s.data.get(), 1, s.size, f);
  std::fwrite(tail.data(), 1, tail.size(), f);
  std::fclose(f);
}

// ---- "Bgeo V5", big-endian ---------------------------------------------------------------------------
struct Bytes {
  std::string s;
  void u8(unsigned v) { s.push_back(static_cast<char>(v)); }
  void u16(unsigned v) { u8(v >> 8); u8(v); }
  void u32(uint32_t v) { u8(v >> 24); u8(v >> 16); u8(v >> 8); u8(v); }
  void str(const char* t) {  // Houdini string: 16-bit length, then the characters
    const size_t n = std::strlen(t);
    u16(static_cast<unsigned>(n));
    s.append(t, n);
  }
};

inline char* put_be32(char* p, uint32_t v) {
  p[0] = static_cast<char>(v >> 24); p[1] = static_cast<char>(v >> 16); p[2] = static_cast<char>(v >> 8); p[3] = static_cast<char>(v);
  return p + 4;
}
inline char* put_be32(char* p, float f) {
  uint32_t v;
  std::memcpy(&v, &f, 4);
  return put_be32(p, v);
}

constexpr size_t kBgeoPointWords = 4 + 3 + 3 + 1 + 1 + 1;  // position xyzw, velocity, color, id, mass, pscale

void write_job_bgeo(const FrameJob& job, int threads) {
  const size_t n = job.points.size();
  Bytes head;
  head.u32(0x4267656fu);  // "Bgeo"
  head.u8('V');
  head.u32(5);
  head.u32(static_cast<uint32_t>(n));  // points
  head.u32(1);                         // primitives
  head.u32(0);                         // point groups
  head.u32(0);                         // primitive groups
  head.u32(5);                         // point attributes (all but position)
  head.u32(0);                         // vertex attributes
  head.u32(1);                         // primitive attributes
  head.u32(0);                         // detail attributes
  // definitions in the order PartioFunctions.h:8-13 adds them; Houdini types: 0 float, 1 int, 5 vector
  const struct { const char* name; unsigned size; uint32_t type; } attrs[5] = {
      {"velocity", 3, 5}, {"color", 3, 5}, {"id", 1, 1}, {"mass", 1, 0}, {"pscale", 1, 0}};
  for (const auto& a : attrs) {
    head.str(a.name);
    head.u16(a.size);
    head.u32(a.type);
    for (unsigned k = 0; k < a.size; ++k) head.u32(0);  // default values
  }

  const bool wide = n > (1u << 16);  // vertex numbers: 32 bits above 65536 points, else 16
  const size_t index_bytes = wide ? 4 : 2;
  std::unique_ptr<char[]> points(new char[n * kBgeoPointWords * 4 + 4]);
  std::unique_ptr<char[]> indices(new char[n * index_bytes + 4]);
  const size_t parts = std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(std::max(1, threads)), (n + 65535) / 65536));
  auto work = [&](size_t k) {
    const size_t i0 = n * k / parts, i1 = n * (k + 1) / parts;
    char* p = points.get() + i0 * kBgeoPointWords * 4;
    char* x = indices.get() + i0 * index_bytes;
    for (size_t i = i0; i < i1; ++i) {
      const FramePoint& q = job.points[i];
      float r, g, b;
      density_colour(q.rho, &r, &g, &b);
      p = put_be32(p, q.px); p = put_be32(p, q.py); p = put_be32(p, q.pz); p = put_be32(p, 1.f);
      p = put_be32(p, q.vx); p = put_be32(p, q.vy); p = put_be32(p, q.vz);
      p = put_be32(p, r); p = put_be32(p, g); p = put_be32(p, b);
      p = put_be32(p, static_cast<uint32_t>(i));
      p = put_be32(p, job.mass);
      p = put_be32(p, job.pscale);
      if (wide) x = put_be32(x, static_cast<uint32_t>(i));
      else { x[0] = static_cast<char>(i >> 8); x[1] = static_cast<char>(i); x += 2; }
    }
  };
  std::vector<std::thread> pool;
  for (size_t k = 1; k < parts; ++k) pool.emplace_back(work, k);
  work(0);
  for (std::thread& t : pool) t.join();

  Bytes mid;  // the primitive attribute table and the head of the one primitive
  mid.str("generator");
  mid.u16(1);   // one value
  mid.u32(4);   // type: index into a string table
  mid.u32(1);   // one string
  mid.str("papi");
  mid.u32(0x8000);  // particle system
  mid.u32(static_cast<uint32_t>(n));
  Bytes tail;
  tail.u32(0);  // the primitive's "generator" value: string 0
  tail.u8(0x00);
  tail.u8(0xff);

  std::FILE* f = std::fopen(job.file_name.c_str(), "wb");
  if (!f) {
    std::cerr << "Error while writing to " << job.file_name << std::endl;
    return;
  }
  std::fwrite(head.s.data(), 1, head.s.size(), f);
  std::fwrite(points.get(), 1, n * kBgeoPointWords * 4, f);
  std::fwrite(mid.s.data(), 1, mid.s.size(), f);
  std::fwrite(indices.get(), 1, n * index_bytes, f);
  std::fwrite(tail.s.data(), 1, tail.s.size(), f);
  std::fclose(f);
}

void write_frame(const FrameJob& job, int threads) {
  if (job.bgeo) write_job_bgeo(job, threads);
  else write_job(job, threads);
}

}  // namespace

// Background writer: one thread taking jobs in order; each job is formatted on `threads` cores.
struct houdini_file_saver::writer {
  std::mutex m;
  std::condition_variable cv;
  std::deque<FrameJob> queue;
  bool busy = false, stop = false;
  int threads = 1;
  std::thread thread;

  writer() {
    const unsigned hw = std::thread::hardware_concurrency();
    threads = static_cast<int>(hw ? hw : 4);
    thread = std::thread([this] { run(); });
  }
  ~writer() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    thread.join();  // run() drains the queue before it leaves
  }
  void run() {
    for (;;) {
      FrameJob job;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return stop || !queue.empty(); });
        if (queue.empty()) return;
        job = std::move(queue.front());
        queue.pop_front();
        busy = true;
      }
      cv.notify_all();
      write_frame(job, threads);
      {
        std::lock_guard<std::mutex> lk(m);
        busy = false;
      }
      cv.notify_all();
    }
  }
  void push(FrameJob&& job) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return queue.size() + (busy ? 1u : 0u) < 2u; });  // at most two frames in flight
    queue.push_back(std::move(job));
    lk.unlock();
    cv.notify_all();
  }
  void drain() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return queue.empty() && !busy; });
  }
};

// For the tests: the writer's number formatting on its own (out needs 32 bytes); returns the length.
extern "C" int clsph_host_format_g(float v, char* out) {
  char* end = put_g(out, v);
  *end = 0;
  return static_cast<int>(end - out);
}

#ifdef USE_PARTIO  // the reference's compile-time switch (houdini_file_saver.cpp:8, :27)
static const houdini_file_saver::frame_format kDefaultFormat = houdini_file_saver::bgeo;
#else
static const houdini_file_saver::frame_format kDefaultFormat = houdini_file_saver::geo;
#endif

houdini_file_saver::houdini_file_saver(std::string prefix)
    : frames_folder_prefix(prefix), asynchronous(true), format(kDefaultFormat), frame_count(0), writer_(nullptr) {}

houdini_file_saver::houdini_file_saver(const houdini_file_saver& other)
    : frames_folder_prefix(other.frames_folder_prefix), asynchronous(other.asynchronous), format(other.format),
      frame_count(other.frame_count), writer_(nullptr) {}

houdini_file_saver& houdini_file_saver::operator=(const houdini_file_saver& other) {
  if (this != &other) {
    wait();
    frames_folder_prefix = other.frames_folder_prefix;
    asynchronous = other.asynchronous;
    format = other.format;
    frame_count = other.frame_count;
  }
  return *this;
}

houdini_file_saver::~houdini_file_saver() { delete writer_; }

void houdini_file_saver::wait() {
  if (writer_) writer_->drain();
}

int houdini_file_saver::writeFrameToFile(particle* particles, const simulation_parameters& parameters) {
  FrameJob job;
  job.bgeo = format == bgeo;
  job.file_name = frames_folder_prefix + "frames/frame" + frame_suffix(++frame_count) + (job.bgeo ? ".bgeo" : ".geo");
  job.mass = parameters.particle_mass;
  job.pscale = parameters.h;
  const unsigned int n = parameters.particles_count;
  job.points.resize(n);
  for (unsigned int i = 0; i < n; ++i) {
    const particle& q = particles[i];
    FramePoint& o = job.points[i];
    o.px = q.position.s[0]; o.py = q.position.s[1]; o.pz = q.position.s[2];
    o.vx = q.velocity.s[0]; o.vy = q.velocity.s[1]; o.vz = q.velocity.s[2];
    o.rho = q.density;
  }
  return submit_job(&job);
}

int houdini_file_saver::writeFramePoints(const float* points, unsigned int count, float particle_mass, float support_radius) {
  static_assert(sizeof(FramePoint) == 7 * sizeof(float), "FramePoint is the packed record of clsph_frame_begin");
  FrameJob job;
  job.bgeo = format == bgeo;
  job.file_name = frames_folder_prefix + "frames/frame" + frame_suffix(++frame_count) + (job.bgeo ? ".bgeo" : ".geo");
  job.mass = particle_mass;
  job.pscale = support_radius;
  job.points.resize(count);
  if (count) std::memcpy(job.points.data(), points, sizeof(FramePoint) * static_cast<size_t>(count));
  return submit_job(&job);
}

int houdini_file_saver::submit_job(void* job_ptr) {
  FrameJob& job = *static_cast<FrameJob*>(job_ptr);
  if (!asynchronous) {
    const unsigned hw = std::thread::hardware_concurrency();
    write_frame(job, static_cast<int>(hw ? hw : 4));
    return 0;
  }
  if (!writer_) writer_ = new writer();
  writer_->push(std::move(job));
  return 0;
}
