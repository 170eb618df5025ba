// TEST INFRASTRUCTURE -- minimal stand-in for the OpenCV C++ API subset used by the reference's
// front-end sources, so that those sources can be compiled UNMODIFIED, from where they lie under
// /root/reference, into oracle/_ref/libplvi_ref.so (recipe: oracle/Makefile.ref).
//
// OpenCV itself (an un-vendored dependency of the reference: headers and libraries are absent in
// this image) is NOT reproduced here: the containers (Mat, Point, KeyPoint, ...) are re-implemented
// from the public OpenCV API documentation, and every image-processing primitive the reference calls
// (cv::FAST, resize, GaussianBlur, copyMakeBorder, pyrDown, Sobel, fastAtan2, LineIterator::count,
// cvRound/cvFloor/cvCeil) forwards to the scalar models of oracle/*.cpp, which are pinned bit-exactly
// against cv2 4.13 (tests/test_oracle_vs_cv2.py, tests/golden/).  So libplvi_ref.so =
// reference-owned logic exactly as written by its authors + restated OpenCV primitives.
//
// Functions the hot path never reaches (drawing, colour conversion, EDLines, ...) are declared so the
// files compile and throw when called.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <fstream>
#include <iomanip>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef signed char schar;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_EXPORTS_W_SIMPLE
#define CV_OUT
#define CV_IN_OUT
#define CV_WRAP
#define CV_PROP
#define CV_PROP_RW
#define CV_PI 3.1415926535897932384626433832795
#define CV_2PI 6.283185307179586476925286766559
#define CV_LOG2 0.69314718055994530941723212145818

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(d, cn) (CV_MAT_DEPTH(d) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_BGR2GRAY 6
#define CV_GRAY2BGR 8

#define CV_Assert(expr) \
  do { if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); } while (0)
#define CV_DbgAssert(expr) ((void)0)
#define CV_Error(code, msg) throw std::runtime_error(std::string("CV_Error: ") + (msg))
#define CV_INSTRUMENT_REGION()

// scalar models in oracle/*.cpp (libplvi_oracle.so)
extern "C" {
float plvio_fast_atan2(float y, float x);
void plvio_resize_linear_u8(const uchar* src, int sstride, int sw, int sh, uchar* dst, int dstride, int dw, int dh);
void plvio_gaussian_blur7_u8(const uchar* src, int sstride, int w, int h, uchar* dst, int dstride);
void plvio_gaussian_blur5_u8(const uchar* src, int sstride, int w, int h, uchar* dst, int dstride);
int plvio_fast_roi(const uchar* img, int stride, int w, int h, int th, float* out, int cap);
void plvio_gaussian_kernel_f64(int n, double sigma, double* k);
void plvio_gaussian_blur_f64(const double* src, int w, int h, double* dst, const double* k, int ksize);
void plvio_resize_linear_f64(const double* src, int sw, int sh, double* dst, int dw, int dh, double fx, double fy);
void plvio_pyr_down_u8(const uchar* src, int w, int h, uchar* dst, int dw, int dh);
void plvio_sobel3_s16(const uchar* src, int w, int h, short* dx, short* dy);
}

// cvRound: round half to even (the SSE2 cvtsd2si / cvtss2si path of OpenCV on x86-64)
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(int v) { return v; }

namespace cv {

using std::max;
using std::min;
typedef std::string String;

[[noreturn]] static inline void cvmini_unreachable(const char* what) {
  throw std::runtime_error(std::string("cvmini: '") + what + "' is outside the hot path and not provided");
}

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

static inline float fastAtan2(float y, float x) { return plvio_fast_atan2(y, x); }

// ------------------------------------------------------------------------------------------
// small geometric types
// ------------------------------------------------------------------------------------------
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
  Point_& operator*=(double s) { x = saturate_cast<T>(x * s); y = saturate_cast<T>(y * s); return *this; }
  Point_& operator*=(float s) { x = saturate_cast<T>(x * s); y = saturate_cast<T>(y * s); return *this; }
  Point_& operator*=(int s) { x = saturate_cast<T>(x * s); y = saturate_cast<T>(y * s); return *this; }
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
  bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
  bool operator!=(const Point_& o) const { return !(*this == o); }
};
template <> template <typename U> inline Point_<float>::Point_(const Point_<U>& o) : x((float)o.x), y((float)o.y) {}
template <> template <typename U> inline Point_<double>::Point_(const Point_<U>& o) : x((double)o.x), y((double)o.y) {}
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
template <typename T> static inline std::ostream& operator<<(std::ostream& os, const Size_<T>& s) { return os << "[" << s.width << " x " << s.height << "]"; }
typedef Size_<int> Size;
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b) : Vec() { val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) : Vec() { val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) : Vec() { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;
typedef Vec<uchar, 3> Vec3b;

struct Scalar {
  double val[4];
  Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar all(double v) { return Scalar(v, v, v, v); }
  double& operator[](int i) { return val[i]; }
  const double& operator[](int i) const { return val[i]; }
  bool operator==(const Scalar& o) const { return !memcmp(val, o.val, sizeof(val)); }
};

struct KeyPoint {   // layout of cv::KeyPoint (28 bytes)
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int im, float d) : queryIdx(q), trainIdx(t), imgIdx(im), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};

template <typename T> struct DataType { enum { type = -1 }; };
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<schar> { enum { type = CV_8SC1 }; };
template <> struct DataType<ushort> { enum { type = CV_16UC1 }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<Vec4f> { enum { type = CV_32FC4 }; };
template <> struct DataType<Vec4i> { enum { type = CV_MAKETYPE(CV_32S, 4) }; };
template <> struct DataType<Vec3b> { enum { type = CV_8UC3 }; };

// ------------------------------------------------------------------------------------------
// Ptr / Algorithm
// ------------------------------------------------------------------------------------------
template <typename T> struct Ptr : public std::shared_ptr<T> {
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
  template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(std::static_pointer_cast<T>(static_cast<const std::shared_ptr<U>&>(o))) {}
  bool empty() const { return !this->get(); }
  void release() { this->reset(); }
  operator T*() const { return this->get(); }
};
template <typename T, typename... A> static inline Ptr<T> makePtr(A&&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

class FileNode {   // yml (de)serialisation is outside the path: every accessor throws
 public:
  template <typename T> void operator>>(T&) const { cvmini_unreachable("FileNode"); }
  FileNode operator[](const char*) const { cvmini_unreachable("FileNode"); }
  FileNode operator[](const String&) const { cvmini_unreachable("FileNode"); }
  FileNode operator[](int) const { cvmini_unreachable("FileNode"); }
  FileNode operator[](unsigned) const { cvmini_unreachable("FileNode"); }
  operator int() const { cvmini_unreachable("FileNode"); }
  operator float() const { cvmini_unreachable("FileNode"); }
  operator double() const { cvmini_unreachable("FileNode"); }
  operator String() const { cvmini_unreachable("FileNode"); }
  size_t size() const { cvmini_unreachable("FileNode"); }
  bool empty() const { return true; }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const String&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const char*) const { cvmini_unreachable("FileStorage"); }
  FileNode operator[](const String&) const { cvmini_unreachable("FileStorage"); }
  template <typename T> FileStorage& operator<<(const T&) { cvmini_unreachable("FileStorage"); }
};
class Algorithm {
 public:
  virtual ~Algorithm() {}
  virtual void clear() {}
  virtual void read(const FileNode&) {}
  virtual void write(FileStorage&) const {}
};

// ------------------------------------------------------------------------------------------
// Mat
// ------------------------------------------------------------------------------------------
static inline size_t cvmini_depth_size(int depth) {
  static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2};
  return s[depth & 7];
}

class _InputArray;
class _OutputArray;
struct MatStep {   // Mat::step converts to size_t and indexes per dimension
  size_t v;
  MatStep() : v(0) {}
  MatStep(size_t s) : v(s) {}
  operator size_t() const { return v; }
  size_t operator[](int i) const { return i == 0 ? v : 0; }
  MatStep& operator=(size_t s) { v = s; return *this; }
};

class Mat {
 public:
  int flags;   // type only
  int dims;
  int rows, cols;
  uchar* data;
  MatStep step;
  std::shared_ptr<std::vector<uchar>> buf;   // owner of the pixels (shared between headers / ROIs)

  Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
  Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); setTo(s); }
  Mat(int r, int c, int type, void* ext, size_t st = 0) : flags(type), dims(2), rows(r), cols(c), data((uchar*)ext) {
    step = st ? st : (size_t)c * elemSize();
  }
  Mat(const Mat& m, const Rect& r) : flags(m.flags), dims(2), rows(r.height), cols(r.width), step(m.step), buf(m.buf) {
    data = m.data + (size_t)r.y * m.step + (size_t)r.x * m.elemSize();
  }
  template <typename T> explicit Mat(const std::vector<T>& v, bool copy = false) : Mat() {
    (void)copy;
    if (!v.empty()) {   // header over the vector's memory (as OpenCV does when copyData == false)
      flags = DataType<T>::type; rows = (int)v.size(); cols = 1; data = (uchar*)v.data(); step = sizeof(T);
    } else {
      flags = DataType<T>::type;
    }
  }

  int type() const { return flags; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize() const { return cvmini_depth_size(depth()) * channels(); }
  size_t elemSize1() const { return cvmini_depth_size(depth()); }
  size_t step1() const { return step / elemSize1(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
  void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }

  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && flags == type) return;   // same geometry: keep the memory (ROI semantics)
    flags = type; rows = r; cols = c; dims = 2;
    step = (size_t)c * elemSize();
    buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 64);
    data = buf->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }

  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
  template <typename T> T* ptr(int r, int c) { return (T*)(data + (size_t)r * step) + c; }
  template <typename T> const T* ptr(int r, int c) const { return (const T*)(data + (size_t)r * step) + c; }
  template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
  template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
  template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
  template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : (cols == 1 ? *(T*)(data + (size_t)i * step) : at<T>(i / cols, i % cols)); }
  template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }

  Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
  Mat row(int i) const { return rowRange(i, i + 1); }
  Mat col(int i) const { return colRange(i, i + 1); }
  Mat operator()(const Rect& r) const { return Mat(*this, r); }

  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& m) const {
    if (empty()) { m.release(); return; }
    m.create(rows, cols, flags);
    const size_t rb = (size_t)cols * elemSize();
    for (int r = 0; r < rows; r++) memmove(m.ptr(r), ptr(r), rb);
  }
  void copyTo(Mat&& m) const { Mat& ref = m; copyTo(ref); }   // dst.row(i) temporaries
  void copyTo(const _OutputArray& o) const;
  void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const;
  void convertTo(const _OutputArray& o, int rtype, double alpha = 1, double beta = 0) const;
  Mat& setTo(const Scalar& s) {
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols * channels(); c++) {
        const double v = s.val[c % channels()];
        switch (depth()) {
          case CV_8U: ptr<uchar>(r)[c] = saturate_cast<uchar>(v); break;
          case CV_16S: ptr<short>(r)[c] = saturate_cast<short>(v); break;
          case CV_32S: ptr<int>(r)[c] = (int)v; break;
          case CV_32F: ptr<float>(r)[c] = (float)v; break;
          case CV_64F: ptr<double>(r)[c] = v; break;
          default: cvmini_unreachable("Mat::setTo depth");
        }
      }
    return *this;
  }
  Mat& operator=(const Scalar& s) { return setTo(s); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); for (int y = 0; y < r; y++) memset(m.ptr(y), 0, (size_t)c * m.elemSize()); return m; }
  static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
  static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar::all(1)); return m; }
  void reserve(size_t) {}
  void push_back(const Mat& m) {   // appends the rows of m (same type and width)
    if (m.empty()) return;
    if (empty()) { m.copyTo(*this); return; }
    CV_Assert(m.type() == type() && m.cols == cols);
    Mat out(rows + m.rows, cols, type());
    const size_t rb = (size_t)cols * elemSize();
    for (int r = 0; r < rows; r++) memcpy(out.ptr(r), ptr(r), rb);
    for (int r = 0; r < m.rows; r++) memcpy(out.ptr(rows + r), m.ptr(r), rb);
    *this = out;
  }
  double dot(const Mat& o) const {   // CV_32F only (pose arithmetic); accumulates in double like cv::Mat::dot
    if (type() != CV_32FC1 || o.type() != CV_32FC1 || rows != o.rows || cols != o.cols) cvmini_unreachable("Mat::dot other than CV_32F");
    double acc = 0;
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) acc += (double)at<float>(r, c) * (double)o.at<float>(r, c);
    return acc;
  }
  Mat t() const {   // CV_32F only (pose arithmetic)
    if (type() != CV_32FC1) cvmini_unreachable("Mat::t other than CV_32F");
    Mat m(cols, rows, CV_32FC1);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
    return m;
  }
  Mat reshape(int cn, int = 0) const {   // continuous matrices only: same memory, another channel count (rows kept)
    if (!isContinuous() || (cols * channels()) % cn) cvmini_unreachable("Mat::reshape of a non-continuous matrix");
    Mat m = *this;
    m.flags = CV_MAKETYPE(depth(), cn);
    m.cols = cols * channels() / cn;
    m.step = (size_t)m.cols * m.elemSize();
    return m;
  }
  Mat inv() const {   // 3x3 CV_32F only (adjugate / determinant in double, rounded once): NOT OpenCV's rounding -- the
                      // tests that reach it invert the identity (unit pinhole K), for which every method is exact
    if (type() != CV_32FC1 || rows != 3 || cols != 3) cvmini_unreachable("Mat::inv other than 3x3 CV_32F");
    double a[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = at<float>(i, j);
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    Mat m(3, 3, CV_32FC1);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        const int r0 = (j + 1) % 3, r1 = (j + 2) % 3, c0 = (i + 1) % 3, c1 = (i + 2) % 3;
        m.at<float>(i, j) = (float)((a[r0][c0] * a[r1][c1] - a[r0][c1] * a[r1][c0]) / det);
      }
    return m;
  }
  static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); if (CV_MAT_DEPTH(type) != CV_32F) cvmini_unreachable("Mat::eye other than CV_32F"); for (int i = 0; i < (r < c ? r : c); i++) m.at<float>(i, i) = 1.0f; return m; }
  int checkVector(int, int = -1, bool = true) const { cvmini_unreachable("Mat::checkVector"); }
};

static inline double cvmini_load(const Mat& m, int r, int c) {
  switch (m.depth()) {
    case CV_8U: return m.ptr<uchar>(r)[c];
    case CV_8S: return m.ptr<schar>(r)[c];
    case CV_16U: return m.ptr<ushort>(r)[c];
    case CV_16S: return m.ptr<short>(r)[c];
    case CV_32S: return m.ptr<int>(r)[c];
    case CV_32F: return m.ptr<float>(r)[c];
    default: return m.ptr<double>(r)[c];
  }
}
inline void Mat::convertTo(Mat& m, int rtype, double alpha, double beta) const {
  const int d = CV_MAT_DEPTH(rtype);
  Mat out(rows, cols, CV_MAKETYPE(d, channels()));
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols * channels(); c++) {
      const double v = cvmini_load(*this, r, c) * alpha + beta;
      switch (d) {
        case CV_8U: out.ptr<uchar>(r)[c] = saturate_cast<uchar>(v); break;
        case CV_16S: out.ptr<short>(r)[c] = saturate_cast<short>(v); break;
        case CV_32S: out.ptr<int>(r)[c] = cvRound(v); break;
        case CV_32F: out.ptr<float>(r)[c] = (float)v; break;
        case CV_64F: out.ptr<double>(r)[c] = v; break;
        default: cvmini_unreachable("Mat::convertTo depth");
      }
    }
  m = out;
}

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() : Mat() { flags = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  explicit Mat_(Size sz) : Mat(sz, DataType<T>::type) {}
  Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { setTo(Scalar::all((double)v)); }
  Mat_(const Mat& m) : Mat() { *this = m; }
  Mat_(const Mat_& m) : Mat(static_cast<const Mat&>(m)) {}
  Mat_& operator=(const Mat_& m) { Mat::operator=(static_cast<const Mat&>(m)); return *this; }
  Mat_& operator=(const Mat& m) {   // converts when the element type differs (cv::Mat_ semantics)
    if (m.empty() || m.type() == (int)DataType<T>::type) Mat::operator=(m);
    else { Mat t; m.convertTo(t, DataType<T>::type); Mat::operator=(t); }
    flags = DataType<T>::type;
    return *this;
  }
  T& operator()(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
  const T& operator()(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
  T& operator()(Point p) { return (*this)(p.y, p.x); }
  T* operator[](int r) { return (T*)(data + (size_t)r * step); }
  const T* operator[](int r) const { return (const T*)(data + (size_t)r * step); }
  using Mat::ptr;
  static Mat_ zeros(Size sz) { return Mat_(Mat::zeros(sz, DataType<T>::type)); }
  static Mat_ zeros(int r, int c) { return Mat_(Mat::zeros(r, c, DataType<T>::type)); }
  Mat_ clone() const { return Mat_(Mat::clone()); }
};

// ------------------------------------------------------------------------------------------
// InputArray / OutputArray proxies (Mat, std::vector<Vec4f>, std::vector<double>, none)
// ------------------------------------------------------------------------------------------
class _InputArray {
 public:
  enum Kind { NONE, MAT, VEC4F, VECD, VECMAT };
  Kind kind;
  void* obj;
  _InputArray() : kind(NONE), obj(nullptr) {}
  _InputArray(const Mat& m) : kind(MAT), obj((void*)&m) {}
  template <typename T> _InputArray(const Mat_<T>& m) : kind(MAT), obj((void*)static_cast<const Mat*>(&m)) {}
  _InputArray(const std::vector<Vec4f>& v) : kind(VEC4F), obj((void*)&v) {}
  _InputArray(const std::vector<double>& v) : kind(VECD), obj((void*)&v) {}
  _InputArray(const std::vector<Mat>& v) : kind(VECMAT), obj((void*)&v) {}
  Mat getMat(int = -1) const {
    switch (kind) {
      case MAT: return *(Mat*)obj;
      case VEC4F: return Mat(*(std::vector<Vec4f>*)obj);
      case VECD: return Mat(*(std::vector<double>*)obj);
      default: return Mat();
    }
  }
  bool empty() const { return kind == NONE || getMat().empty(); }
  int channels() const { return getMat().channels(); }
  int type() const { return getMat().type(); }
  int depth() const { return getMat().depth(); }
  Size size() const { return getMat().size(); }
  bool needed() const { return kind != NONE; }
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) : _InputArray(m) {}
  template <typename T> _OutputArray(Mat_<T>& m) : _InputArray(m) {}
  _OutputArray(std::vector<Vec4f>& v) : _InputArray(v) {}
  _OutputArray(std::vector<double>& v) : _InputArray(v) {}
  _OutputArray(std::vector<Mat>& v) : _InputArray(v) {}
  Mat& getMatRef() const { if (kind != MAT) cvmini_unreachable("getMatRef on non-Mat"); return *(Mat*)obj; }
  void create(int r, int c, int type) const {
    switch (kind) {
      case MAT: ((Mat*)obj)->create(r, c, type); break;
      case VEC4F: ((std::vector<Vec4f>*)obj)->resize((size_t)r * c); break;
      case VECD: ((std::vector<double>*)obj)->resize((size_t)r * c); break;
      default: cvmini_unreachable("OutputArray::create on noArray");
    }
  }
  void create(Size sz, int type) const { create(sz.height, sz.width, type); }
  void release() const {
    switch (kind) {
      case MAT: ((Mat*)obj)->release(); break;
      case VEC4F: ((std::vector<Vec4f>*)obj)->clear(); break;
      case VECD: ((std::vector<double>*)obj)->clear(); break;
      default: break;
    }
  }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
typedef InputArray InputArrayOfArrays;
typedef OutputArray OutputArrayOfArrays;
static inline const _OutputArray& noArray() { static _OutputArray none; return none; }

inline void Mat::copyTo(const _OutputArray& o) const {
  if (o.kind == _InputArray::MAT) { copyTo(o.getMatRef()); return; }
  if (empty()) { o.release(); return; }
  o.create(rows, cols, flags);
  Mat dst = o.getMat();
  if (dst.elemSize() != elemSize()) cvmini_unreachable("copyTo(vector) with another element type");
  for (int r = 0; r < rows; r++) memcpy(dst.data + (size_t)r * cols * elemSize(), ptr(r), (size_t)cols * elemSize());
}
inline void Mat::convertTo(const _OutputArray& o, int rtype, double alpha, double beta) const {
  convertTo(o.getMatRef(), rtype, alpha, beta);
}

// ------------------------------------------------------------------------------------------
// image-processing primitives -> oracle scalar models
// ------------------------------------------------------------------------------------------
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_RGB2GRAY = 7 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { LINE_8 = 8, LINE_AA = 16 };

static inline Mat cvmini_packed(const Mat& m) { return m.isContinuous() ? m : m.clone(); }

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
static inline int cvmini_reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): TYPE_9_16, KeyPoint(x, y, 7.f, -1, score)
static inline void FAST(InputArray image, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
  if (!nms) cvmini_unreachable("FAST without non-maximum suppression");
  const Mat m = image.getMat();
  CV_Assert(m.type() == CV_8UC1);
  kps.clear();
  std::vector<float> out((size_t)3 * m.rows * m.cols + 3);
  const int n = plvio_fast_roi(m.data, (int)m.step, m.cols, m.rows, threshold, out.data(), m.rows * m.cols);
  for (int i = 0; i < n; i++) kps.push_back(KeyPoint(out[3 * i], out[3 * i + 1], 7.f, -1.f, out[3 * i + 2]));
}

static inline void resize(InputArray src_, OutputArray dst_, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR) {
  if (interp != INTER_LINEAR) cvmini_unreachable("resize with an interpolation other than INTER_LINEAR");
  const Mat src = src_.getMat();
  if (dsize.area() == 0) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
  else { fx = (double)dsize.width / src.cols; fy = (double)dsize.height / src.rows; }
  if (src.type() == CV_8UC1) {
    const Mat s = src.clone();                // the destination may alias the source's parent buffer
    dst_.create(dsize, src.type());
    Mat dst = dst_.getMat();
    plvio_resize_linear_u8(s.data, (int)s.step, s.cols, s.rows, dst.data, (int)dst.step, dsize.width, dsize.height);
  } else if (src.type() == CV_64FC1) {
    const Mat s = src.clone();
    Mat tmp(dsize, CV_64FC1);
    plvio_resize_linear_f64(s.ptr<double>(), s.cols, s.rows, tmp.ptr<double>(), dsize.width, dsize.height, fx, fy);
    dst_.create(dsize, src.type());
    tmp.copyTo(dst_.getMatRef());
  } else {
    cvmini_unreachable("resize of this element type");
  }
}

static inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY = 0,
                                int borderType = BORDER_DEFAULT) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) cvmini_unreachable("GaussianBlur border type");
  if (sigmaY == 0) sigmaY = sigmaX;
  if (ksize.width != ksize.height || sigmaX != sigmaY) cvmini_unreachable("anisotropic GaussianBlur");
  const Mat s = src_.getMat().clone();        // in-place calls; ROI sources are blurred as isolated images
  if (s.type() == CV_8UC1) {
    Mat tmp(s.rows, s.cols, CV_8UC1);
    if (ksize.width == 7 && sigmaX == 2) plvio_gaussian_blur7_u8(s.data, (int)s.step, s.cols, s.rows, tmp.data, (int)tmp.step);
    else if (ksize.width == 5 && sigmaX == 1) plvio_gaussian_blur5_u8(s.data, (int)s.step, s.cols, s.rows, tmp.data, (int)tmp.step);
    else cvmini_unreachable("8-bit GaussianBlur with this kernel");
    dst_.create(s.rows, s.cols, CV_8UC1);
    tmp.copyTo(dst_.getMatRef());
  } else if (s.type() == CV_64FC1) {
    std::vector<double> k(ksize.width);
    plvio_gaussian_kernel_f64(ksize.width, sigmaX, k.data());
    Mat tmp(s.rows, s.cols, CV_64FC1);
    plvio_gaussian_blur_f64(s.ptr<double>(), s.cols, s.rows, tmp.ptr<double>(), k.data(), ksize.width);
    dst_.create(s.rows, s.cols, CV_64FC1);
    tmp.copyTo(dst_.getMatRef());
  } else {
    cvmini_unreachable("GaussianBlur of this element type");
  }
}

static inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType,
                                  const Scalar& = Scalar()) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) cvmini_unreachable("copyMakeBorder border type");
  // Without BORDER_ISOLATED OpenCV reads real pixels outside a ROI; the reference's only such call passes a whole image.
  const Mat s = src_.getMat().clone();
  CV_Assert(s.elemSize() == 1);
  dst_.create(s.rows + top + bottom, s.cols + left + right, s.type());
  Mat dst = dst_.getMat();
  for (int y = 0; y < dst.rows; y++) {
    const uchar* sr = s.ptr(cvmini_reflect101(y - top, s.rows));
    uchar* dr = dst.ptr(y);
    for (int x = 0; x < dst.cols; x++) dr[x] = sr[cvmini_reflect101(x - left, s.cols)];
  }
}

static inline void pyrDown(InputArray src_, OutputArray dst_, const Size& dstsize = Size(), int = BORDER_DEFAULT) {
  const Mat s = cvmini_packed(src_.getMat()).clone();
  CV_Assert(s.type() == CV_8UC1);
  const Size ds = dstsize.area() == 0 ? Size((s.cols + 1) / 2, (s.rows + 1) / 2) : dstsize;
  Mat tmp(ds, CV_8UC1);
  plvio_pyr_down_u8(s.data, s.cols, s.rows, tmp.data, ds.width, ds.height);
  dst_.create(ds, CV_8UC1);
  tmp.copyTo(dst_.getMatRef());
}

static inline void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize = 3, double scale = 1,
                         double delta = 0, int = BORDER_DEFAULT) {
  const Mat s = cvmini_packed(src_.getMat()).clone();
  if (s.type() != CV_8UC1 || ddepth != CV_16S || ksize != 3 || scale != 1 || delta != 0 || dx + dy != 1)
    cvmini_unreachable("Sobel with these arguments");
  Mat gx(s.rows, s.cols, CV_16SC1), gy(s.rows, s.cols, CV_16SC1);
  plvio_sobel3_s16(s.data, s.cols, s.rows, gx.ptr<short>(), gy.ptr<short>());
  dst_.create(s.rows, s.cols, CV_16SC1);
  (dx ? gx : gy).copyTo(dst_.getMatRef());
}

// cv::clipLine(Size, Point&, Point&) (imgproc/src/drawing.cpp): integer Cohen-Sutherland clip to
// [0, w-1] x [0, h-1]; restated in oracle_line.cpp (plvio_clip_line) and pinned against cv2.clipLine.
extern "C" int plvio_clip_line(int w, int h, int* x1, int* y1, int* x2, int* y2);
static inline bool clipLine(Size sz, Point& p1, Point& p2) { return plvio_clip_line(sz.width, sz.height, &p1.x, &p1.y, &p2.x, &p2.y) != 0; }

// cv::LineIterator(img, pt1, pt2, 8): only `count` is consumed by the reference (LSDDetector_custom.cpp:333-334).
// Point2f arguments convert to Point with cvRound (saturate_cast<int>); endpoints outside the image are clipped
// first; 8-connected count = max(|dx|, |dy|) + 1, or 0 when the line misses the image.
class LineIterator {
 public:
  int count;
  LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity = 8, bool leftToRight = false) {
    (void)leftToRight;
    if (connectivity != 8) cvmini_unreachable("LineIterator connectivity");
    count = -1;
    if ((unsigned)pt1.x >= (unsigned)img.cols || (unsigned)pt2.x >= (unsigned)img.cols ||
        (unsigned)pt1.y >= (unsigned)img.rows || (unsigned)pt2.y >= (unsigned)img.rows) {
      if (!clipLine(img.size(), pt1, pt2)) { count = 0; return; }
    }
    count = std::max(std::abs(pt2.x - pt1.x), std::abs(pt2.y - pt1.y)) + 1;
  }
};

class KeyPointsFilter {
 public:
  static void retainBest(std::vector<KeyPoint>&, int) { cvmini_unreachable("KeyPointsFilter::retainBest"); }
};

class LineSegmentDetector : public Algorithm {
 public:
  virtual void detect(InputArray image, OutputArray lines, OutputArray width = noArray(), OutputArray prec = noArray(),
                      OutputArray nfa = noArray()) = 0;
  virtual void drawSegments(InputOutputArray image, InputArray lines) = 0;
  virtual int compareSegments(const Size& size, InputArray lines1, InputArray lines2, InputOutputArray image = noArray()) = 0;
  virtual ~LineSegmentDetector() {}
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8, double sigma_scale = 0.6,
                                                   double quant = 2.0, double ang_th = 22.5, double log_eps = 0,
                                                   double density_th = 0.7, int n_bins = 1024);

// outside the hot path: declared, never reached
template <typename... A> static inline void cvtColor(A&&...) { cvmini_unreachable("cvtColor"); }
template <typename... A> static inline void line(A&&...) { cvmini_unreachable("line"); }
template <typename... A> static inline void split(A&&...) { cvmini_unreachable("split"); }
template <typename... A> static inline void merge(A&&...) { cvmini_unreachable("merge"); }
template <typename... A> static inline void bitwise_xor(A&&...) { cvmini_unreachable("bitwise_xor"); }
template <typename... A> static inline void bitwise_and(A&&...) { cvmini_unreachable("bitwise_and"); }
template <typename... A> static inline void bitwise_or(A&&...) { cvmini_unreachable("bitwise_or"); }
template <typename... A> static inline void bitwise_not(A&&...) { cvmini_unreachable("bitwise_not"); }
template <typename... A> static inline int countNonZero(A&&...) { cvmini_unreachable("countNonZero"); }
template <typename... A> static inline void imshow(A&&...) { cvmini_unreachable("imshow"); }
template <typename... A> static inline int waitKey(A&&...) { cvmini_unreachable("waitKey"); }
template <typename... A> static inline void circle(A&&...) { cvmini_unreachable("circle"); }
template <typename... A> static inline void rectangle(A&&...) { cvmini_unreachable("rectangle"); }
template <typename... A> static inline void putText(A&&...) { cvmini_unreachable("putText"); }
template <typename... A> static inline void add(A&&...) { cvmini_unreachable("add"); }
template <typename... A> static inline double threshold(A&&...) { cvmini_unreachable("threshold"); }
template <typename... A> static inline void compare(A&&...) { cvmini_unreachable("compare"); }
enum { THRESH_BINARY = 0, THRESH_TOZERO = 3 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
static inline Mat abs(const Mat&) { cvmini_unreachable("abs(Mat)"); }
static inline Mat operator/(const Mat& a, double d) {   // CV_32F only (pose arithmetic): element / d, rounded to float once
  if (a.type() != CV_32FC1) cvmini_unreachable("Mat / scalar other than CV_32F");
  Mat c(a.rows, a.cols, CV_32FC1);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = (float)((double)a.at<float>(i, j) / d);
  return c;
}
static inline Mat operator*(double d, const Mat& a) {   // CV_32F only (pose arithmetic): d * element, rounded to float once
  if (a.type() != CV_32FC1) cvmini_unreachable("scalar * Mat other than CV_32F");
  Mat c(a.rows, a.cols, CV_32FC1);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = (float)(d * (double)a.at<float>(i, j));
  return c;
}
static inline double norm(const Mat& a, int t = NORM_L2) {   // CV_32F L2 only (pose arithmetic), double accumulation
  if (a.type() != CV_32FC1 || t != NORM_L2) cvmini_unreachable("norm(Mat) other than CV_32F L2");
  return std::sqrt(a.dot(a));
}
static inline double norm(const Mat& a, const Mat& b, int t = NORM_L2) {   // CV_16S NORM_L1 only (Frame::ComputeStereoMatches' SAD)
  if (a.type() != CV_16SC1 || b.type() != CV_16SC1 || t != NORM_L1 || a.rows != b.rows || a.cols != b.cols)
    cvmini_unreachable("norm(Mat, Mat) other than CV_16S NORM_L1");
  double s = 0;   // cv::norm accumulates the L1 norm of 16-bit data in int and returns it as double: exact either way here
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += std::abs((int)a.at<short>(r, c) - (int)b.at<short>(r, c));
  return s;
}
static inline Mat operator-(const Mat& a, short v) {   // MatExpr a - Scalar on CV_16S: saturate_cast<short>(a - v)
  if (a.type() != CV_16SC1) cvmini_unreachable("Mat - scalar other than CV_16S");
  Mat c(a.rows, a.cols, CV_16SC1);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) c.at<short>(i, j) = saturate_cast<short>((int)a.at<short>(i, j) - (int)v);
  return c;
}
// cv::undistortPoints(src, dst, K, dist, noArray(), P): N x 1 CV_32FC2 points, CV_32F 3x3 K / P, 4..14 CV_32F distortion
// coefficients -> the oracle's restatement of OpenCV 4.x (pinned against cv2.undistortPoints, tests/test_frame_cpu.py)
extern "C" void plvio_cv_undistort_points(const float* xy, int n, const double* K, const double* k, const double* P, float* out);
static inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R = Mat(), const Mat& P = Mat()) {
  if (src.type() != CV_MAKETYPE(CV_32F, 2) || !src.isContinuous() || K.type() != CV_32FC1 || dist.type() != CV_32FC1 || !R.empty() ||
      P.type() != CV_32FC1)
    cvmini_unreachable("undistortPoints other than CV_32FC2 points with CV_32F K / dist / P and no R");
  const double k4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
  const double p4[4] = {P.at<float>(0, 0), P.at<float>(1, 1), P.at<float>(0, 2), P.at<float>(1, 2)};
  double k[14] = {0};
  const int nk = (int)dist.total();
  for (int i = 0; i < nk && i < 14; i++) k[i] = dist.at<float>(i);
  const int n = (int)src.total();
  std::vector<float> out(2 * (size_t)n + 2);
  plvio_cv_undistort_points(src.ptr<float>(), n, k4, k, p4, out.data());
  Mat d(src.rows, src.cols, src.type());
  memcpy(d.data, out.data(), sizeof(float) * 2 * (size_t)n);
  dst = d;
}
static inline void hconcat(const Mat&, const Mat&, Mat&) { cvmini_unreachable("hconcat"); }
static inline void vconcat(const Mat&, const Mat&, Mat&) { cvmini_unreachable("vconcat"); }
template <typename T> struct MatCommaInitializer_ {   // (Mat_<T>(r, c) << a, b, ...)
  Mat_<T> m;
  int i;
  MatCommaInitializer_(const Mat_<T>& m_, T first) : m(m_), i(0) { *this, first; }
  MatCommaInitializer_& operator,(T v) { if (i < (int)m.total()) m(i / m.cols, i % m.cols) = v; i++; return *this; }
  operator Mat_<T>() const { return m; }
  operator Mat() const { return m; }
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) { return MatCommaInitializer_<T>(m, (T)v); }

// cv::BFMatcher(NORM_HAMMING).knnMatch: the k nearest train rows of every query row, ordered by (distance, train
// index) -- the tie rule probed on cv2.BFMatcher (tests/test_oracle_vs_cv2.py); fewer than k entries when the train
// set is smaller.
class BFMatcher : public Algorithm {
 public:
  BFMatcher(int normType = NORM_L2, bool crossCheck = false) {
    if (normType != NORM_HAMMING || crossCheck) cvmini_unreachable("BFMatcher other than NORM_HAMMING without cross-check");
  }
  static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false) { return Ptr<BFMatcher>(new BFMatcher(normType, crossCheck)); }
  void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& matches, int k) const {
    CV_Assert(q.type() == CV_8UC1 && (t.empty() || (t.type() == CV_8UC1 && t.cols == q.cols)));
    matches.assign(q.rows, std::vector<DMatch>());
    for (int i = 0; i < q.rows; i++) {
      std::vector<std::pair<int, int>> d(t.rows);
      for (int j = 0; j < t.rows; j++) {
        int c = 0;
        for (int b = 0; b < q.cols; b++) c += __builtin_popcount((unsigned)(q.ptr(i)[b] ^ t.ptr(j)[b]));
        d[j] = std::make_pair(c, j);
      }
      const int kk = std::min(k, t.rows);
      std::partial_sort(d.begin(), d.begin() + kk, d.end());
      for (int r = 0; r < kk; r++) matches[i].push_back(DMatch(i, d[r].second, (float)d[r].first));
    }
  }
};
// Small CV_32F matrix algebra (pose arithmetic of the matchers: 3x3 / 3x1 / 4x4).  Products accumulate in double and
// round once per element.  cv::gemm's rounding for such sizes is NOT modelled: the tests that reach these operators use
// poses whose products are exact in any order (identity rotations, integer-valued terms), see slam_mock_orb.h.
static inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.type() != CV_32FC1 || b.type() != CV_32FC1 || a.cols != b.rows) cvmini_unreachable("Mat * Mat other than CV_32F with matching sizes");
  Mat c(a.rows, b.cols, CV_32FC1);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double acc = 0;
      for (int k = 0; k < a.cols; k++) acc += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
      c.at<float>(i, j) = (float)acc;
    }
  return c;
}
static inline Mat cvmini_f32_elementwise(const Mat& a, const Mat* b, int op, const char* what) {
  if (a.type() != CV_32FC1 || (b && (b->type() != CV_32FC1 || b->rows != a.rows || b->cols != a.cols))) cvmini_unreachable(what);
  Mat c(a.rows, a.cols, CV_32FC1);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      const float x = a.at<float>(i, j), y = b ? b->at<float>(i, j) : 0.f;
      c.at<float>(i, j) = op == 0 ? x + y : (op == 1 ? x - y : -x);
    }
  return c;
}
static inline Mat operator*(const Mat&, double) { cvmini_unreachable("Mat * scalar"); }
static inline Mat operator+(const Mat& a, const Mat& b) { return cvmini_f32_elementwise(a, &b, 0, "Mat + Mat other than CV_32F"); }
static inline Mat operator-(const Mat& a, const Mat& b) { return cvmini_f32_elementwise(a, &b, 1, "Mat - Mat other than CV_32F"); }
static inline Mat operator-(const Mat& a) { return cvmini_f32_elementwise(a, nullptr, 2, "-Mat other than CV_32F"); }

}  // namespace cv
