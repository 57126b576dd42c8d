// MOCK of the few gtsam / Eigen types integration/ros/IMUManager_vlo.cpp touches (compile + link check only: the image has
// neither Eigen nor GTSAM).  Names, member names and signatures follow GTSAM 4.0.x with manifold preintegration
// (GTSAM_TANGENT_PREINTEGRATION off): gtsam/navigation/ManifoldPreintegration.h, CombinedImuFactor.h.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <boost_shim.h>
namespace gtsam {
template <int N> struct VecN {
  std::array<double, N> v{};
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  static VecN Zero() { return VecN(); }
  const double *data() const { return v.data(); }
  double *data() { return v.data(); }
};
using Vector3 = VecN<3>;
using Vector6 = VecN<6>;
inline Vector3 operator*(double s, const Vector3 &a) { Vector3 r; for (int i = 0; i < 3; i++) r[i] = s * a[i]; return r; }
inline Vector3 operator+(const Vector3 &a, const Vector3 &b) { Vector3 r; for (int i = 0; i < 3; i++) r[i] = a[i] + b[i]; return r; }
template <int R, int C> struct MatRC {
  std::array<double, R * C> m{};      // row-major in this mock
  double &operator()(int r, int c) { return m[r * C + c]; }
  double operator()(int r, int c) const { return m[r * C + c]; }
};
using Matrix3 = MatRC<3, 3>;
using Matrix15 = MatRC<15, 15>;
}  // namespace gtsam
