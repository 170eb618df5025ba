// TEST INFRASTRUCTURE -- stand-in for the three Eigen operations src/LineExtractor.cc:106-116 uses on
// Eigen::Vector3d (comma initialiser, cross product, division by a scalar); Eigen is absent in this image.
// See cvmini.hpp for the purpose (compiling the reference's sources unmodified into oracle/_ref).
#pragma once
namespace Eigen {
struct Vector2d {   // named by Frame.h / Frame.cc (stereo-line helpers, never reached)
  double v[2];
  Vector2d() { v[0] = v[1] = 0; }
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
};
struct Vector3d {
  double v[3];
  Vector3d() { v[0] = v[1] = v[2] = 0; }
  Vector3d(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  struct CommaInit {
    Vector3d& t; int n;
    CommaInit& operator,(double x) { t.v[n++] = x; return *this; }
  };
  CommaInit operator<<(double x) { v[0] = x; return CommaInit{*this, 1}; }
  CommaInit operator<<(const Vector3d& o) { *this = o; return CommaInit{*this, 3}; }
  Vector3d cross(const Vector3d& o) const {   // same expression order as Eigen's cross3 (a1*b2 - a2*b1, ...)
    return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  Vector3d operator/(double s) const { return Vector3d(v[0] / s, v[1] / s, v[2] / s); }
  Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vector2d head(int) const { Vector2d r; r.v[0] = v[0]; r.v[1] = v[1]; return r; }
};
#ifdef SLAM_MOCK_REAL_MAPLINE
// include/MapLine.h:34-35 names Eigen::Matrix<double,6,6> / <double,6,1>; MapLine.cc assigns to head(3) / tail(3)
template <typename T, int R, int C> struct Matrix {
  T v[R * C];
  Matrix() { for (T& x : v) x = 0; }
  struct Seg {
    T* p;
    Seg& operator=(const Vector3d& o) { p[0] = o.v[0]; p[1] = o.v[1]; p[2] = o.v[2]; return *this; }
    operator Vector3d() const { return Vector3d(p[0], p[1], p[2]); }
  };
  Seg head(int) { return Seg{v}; }
  Seg tail(int n) { return Seg{v + R * C - n}; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
};
typedef Matrix<double, 6, 1> Vector6d;
#else
struct Vector6d {   // Eigen::Matrix<double, 6, 1>: only head(3) / tail(3) / operator() are used (never reached)
  double v[6];
  Vector6d() { for (double& x : v) x = 0; }
  Vector3d head(int) const { return Vector3d(v[0], v[1], v[2]); }
  Vector3d tail(int) const { return Vector3d(v[3], v[4], v[5]); }
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
};
#endif
}  // namespace Eigen
