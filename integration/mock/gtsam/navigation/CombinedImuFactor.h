// MOCK (see gtsam/base/Matrix.h): member names as in GTSAM 4.0.x PreintegrationBase / ManifoldPreintegration /
// PreintegratedCombinedMeasurements, so that a subclass adopting a preintegrated POD compiles here as it does there.
#pragma once
#include <gtsam/base/Matrix.h>
#include <gtsam/inference/Symbol.h>
namespace gtsam {
struct Rot3 { Matrix3 R; Rot3() { R(0, 0) = R(1, 1) = R(2, 2) = 1; } explicit Rot3(const Matrix3 &m) : R(m) {} };
using Point3 = Vector3;
struct NavState { Rot3 R_; Point3 t_; Vector3 v_; NavState() {} NavState(const Rot3 &R, const Point3 &t, const Vector3 &v) : R_(R), t_(t), v_(v) {} };
namespace imuBias { struct ConstantBias {
  Vector3 ba, bg; ConstantBias() {} explicit ConstantBias(const Vector6 &b) { for (int i = 0; i < 3; i++) { ba[i] = b[i]; bg[i] = b[3 + i]; } }
  Vector6 vector() const { Vector6 r; for (int i = 0; i < 3; i++) { r[i] = ba[i]; r[3 + i] = bg[i]; } return r; }
}; }
struct PreintegrationCombinedParams { Vector3 n_gravity; };
class PreintegrationBase {
 public:
  typedef imuBias::ConstantBias Bias;
 protected:
  boost::shared_ptr<PreintegrationCombinedParams> p_;
  Bias biasHat_;
  double deltaTij_ = 0;
  PreintegrationBase(const boost::shared_ptr<PreintegrationCombinedParams> &p, const Bias &b) : p_(p), biasHat_(b) {}
 public:
  double deltaTij() const { return deltaTij_; }
};
class ManifoldPreintegration : public PreintegrationBase {
 protected:
  NavState deltaXij_;
  Matrix3 delRdelBiasOmega_, delPdelBiasAcc_, delPdelBiasOmega_, delVdelBiasAcc_, delVdelBiasOmega_;
  ManifoldPreintegration(const boost::shared_ptr<PreintegrationCombinedParams> &p, const Bias &b) : PreintegrationBase(p, b) {}
 public:
  const NavState &deltaXij() const { return deltaXij_; }
};
class PreintegratedCombinedMeasurements : public ManifoldPreintegration {
 public:
  typedef PreintegrationCombinedParams Params;
  PreintegratedCombinedMeasurements(const boost::shared_ptr<Params> &p, const imuBias::ConstantBias &b = imuBias::ConstantBias())
      : ManifoldPreintegration(p, b) {}
  void resetIntegration() { deltaTij_ = 0; deltaXij_ = NavState(); }
  void resetIntegrationAndSetBias(const Bias &b) { biasHat_ = b; resetIntegration(); }
  void integrateMeasurement(const Vector3 &, const Vector3 &, double dt) { deltaTij_ += dt; }
  const Matrix15 &preintMeasCov() const { return preintMeasCov_; }
  const boost::shared_ptr<Params> &params() const { return p_; }
 protected:
  Matrix15 preintMeasCov_;
};
class CombinedImuFactor {
 public:
  CombinedImuFactor(Key, Key, Key, Key, Key, Key, const PreintegratedCombinedMeasurements &pim) : pim_(pim) {}
  const PreintegratedCombinedMeasurements &preintegratedMeasurements() const { return pim_; }
 private:
  PreintegratedCombinedMeasurements pim_;
};
}  // namespace gtsam
