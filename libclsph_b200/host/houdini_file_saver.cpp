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
// bytes) into a job and returns; a background thread formats the job on all host cores
// (std::to_chars, which is specified to match printf %g and is several times faster than
// snprintf) and writes the file with one call. At most two frames are in flight; wait() or the
// destructor blocks until everything is on disk. `asynchronous = false` restores the reference's
// "written when the call returns".
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

// printf("%g", (double)v) into p; returns the end. Non-finite values keep the C library's spelling.
inline char* put_g(char* p, float v) {
  const double d = static_cast<double>(v);
  if (!std::isfinite(d)) return p + std::snprintf(p, 32, "%g", d);
  return std::to_chars(p, p + 32, d, std::chars_format::general, 6).ptr;
}
inline char* put_uint(char* p, unsigned int v) { return std::to_chars(p, p + 16, v).ptr; }

struct FramePoint {  // what a frame needs of a particle: 28 of its 80 bytes
  float px, py, pz, vx, vy, vz, rho;
};

struct FrameJob {
  std::string file_name;
  float mass = 0.f;
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
      write_job(job, threads);
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

houdini_file_saver::houdini_file_saver(std::string prefix)
    : frames_folder_prefix(prefix), asynchronous(true), frame_count(0), writer_(nullptr) {}

houdini_file_saver::houdini_file_saver(const houdini_file_saver& other)
    : frames_folder_prefix(other.frames_folder_prefix), asynchronous(other.asynchronous), frame_count(other.frame_count),
      writer_(nullptr) {}

houdini_file_saver& houdini_file_saver::operator=(const houdini_file_saver& other) {
  if (this != &other) {
    wait();
    frames_folder_prefix = other.frames_folder_prefix;
    asynchronous = other.asynchronous;
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
  job.file_name = frames_folder_prefix + "frames/frame" + frame_suffix(++frame_count) + ".geo";
  job.mass = parameters.particle_mass;
  const unsigned int n = parameters.particles_count;
  job.points.resize(n);
  for (unsigned int i = 0; i < n; ++i) {
    const particle& q = particles[i];
    FramePoint& o = job.points[i];
    o.px = q.position.s[0]; o.py = q.position.s[1]; o.pz = q.position.s[2];
    o.vx = q.velocity.s[0]; o.vy = q.velocity.s[1]; o.vz = q.velocity.s[2];
    o.rho = q.density;
  }
  if (!asynchronous) {
    const unsigned hw = std::thread::hardware_concurrency();
    write_job(job, static_cast<int>(hw ? hw : 4));
    return 0;
  }
  if (!writer_) writer_ = new writer();
  writer_->push(std::move(job));
  return 0;
}
