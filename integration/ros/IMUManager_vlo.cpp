// IMUManager_vlo.cpp -- LINK-TIME replacement of gtsam_fusion/src/gtsam_fusion/IMUManager.cpp.
//
// GraphManager holds a std::shared_ptr<ImuManagerRos> (gtsam_fusion/include/gtsam_fusion/GraphManager.h:101) and calls the
// NON-virtual IMUManager::getFactor (IMUManager.h:26-27) from reserveNode (GraphManager.cpp:51-69): a subclass that shadows
// the method is never dispatched.  What does bind is this translation unit compiled INSTEAD of IMUManager.cpp -- it defines
// the very same members (constructor, addIMUMeasurement, both getFactor overloads) against the reference's unchanged
// header, keeps the buffer / window semantics of IMUManager.cpp:19-79 on the host (samples with t <= t0 are dropped, the
// samples with t < t1 are consumed, the first sample with t >= t1 stays for the interpolated last step and the next call)
// and hands the preintegration itself (gtsam::PreintegratedCombinedMeasurements::integrateMeasurement, IMUManager.cpp:50-52,64)
// to the B200 through vlo_imu_preintegrate_batch.  In CMakeLists.txt: replace src/gtsam_fusion/IMUManager.cpp by this file
// and link libvlo.so.  GTSAM must be built with manifold preintegration (GTSAM_TANGENT_PREINTEGRATION off), whose
// protected members the Adopt class below fills.
#include <gtsam_fusion/IMUManager.h>
#include <gtsam/inference/Symbol.h>
#include <stdexcept>
#include <vector>
#include "vlo.h"

namespace VILFusion
{
    using gtsam::symbol_shorthand::X;  // Pose
    using gtsam::symbol_shorthand::V;  // Velocity
    using gtsam::symbol_shorthand::B;  // IMU Bias

    namespace {
    // one handle per process for the IMU seam (the LiDAR node owns its own); created on first use
    vlo_handle *imu_handle()
    {
        static vlo_handle *h = nullptr;
        if (!h) {
            vlo_config cfg;
            vlo_default_config(&cfg);
            cfg.max_scans = 2; cfg.max_points = 64;                  // no LiDAR work on this handle
            if (vlo_create(&cfg, &h) != VLO_OK) throw std::runtime_error("IMUManager (vlo): no usable CUDA device");
        }
        return h;
    }

    // gtsam keeps the preintegrated state protected: adopt the POD through a thin subclass
    struct Adopt : gtsam::PreintegratedCombinedMeasurements {
        Adopt(const gtsam::PreintegratedCombinedMeasurements &like, const gtsam::imuBias::ConstantBias &bias, const vlo_preint &f)
            : gtsam::PreintegratedCombinedMeasurements(like)
        {
            auto m3 = [](const double *s) { gtsam::Matrix3 M; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M(r, c) = s[r * 3 + c]; return M; };
            biasHat_ = bias;
            deltaTij_ = f.dt;
            gtsam::Vector3 dP, dV;
            for (int k = 0; k < 3; k++) { dP[k] = f.dP[k]; dV[k] = f.dV[k]; }
            deltaXij_ = gtsam::NavState(gtsam::Rot3(m3(f.dR)), dP, dV);
            delRdelBiasOmega_ = m3(f.dR_dbg);
            delPdelBiasAcc_ = m3(f.dP_dba); delPdelBiasOmega_ = m3(f.dP_dbg);
            delVdelBiasAcc_ = m3(f.dV_dba); delVdelBiasOmega_ = m3(f.dV_dbg);
            for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) preintMeasCov_(r, c) = f.cov[r * 15 + c];   // (theta, p, v, ba, bg)
        }
    };
    }  // namespace

    IMUManager::IMUManager(boost::shared_ptr<PreintegratedCombinedMeasurements::Params> imuParams) :
        _integrator(imuParams, imuBias::ConstantBias(Vector6::Zero()))
    {
        _integrator.resetIntegration();
    }

    void IMUManager::addIMUMeasurement(double time, const Vector3 accel, const Vector3 gyro)
    {
        LockGuard guard(_bufferMutex);
        _buffer.push_back(Measurement { time, accel, gyro });
    }

    CombinedImuFactor IMUManager::getFactor(double startTime, double endTime, uint64_t currentIndex, imuBias::ConstantBias bias)
    {
        LockGuard _lockGuard(_bufferMutex);
        // the window the CPU loop walks: every buffered sample up to and including the first one with t >= endTime
        std::vector<double> t, a, w;
        for (const Measurement &m : _buffer) {
            t.push_back(m.time);
            for (int k = 0; k < 3; k++) { a.push_back(m.accel[k]); w.push_back(m.gyro[k]); }
            if (m.time >= endTime) break;
        }
        const Vector6 b = bias.vector();
        double b6[6];
        for (int k = 0; k < 6; k++) b6[k] = b[k];
        vlo_preint f;
        const int rc = vlo_imu_preintegrate_batch(imu_handle(), t.data(), a.data(), w.data(), (int)t.size(), &startTime, &endTime, b6, 1, &f);
        if (rc < 0) throw std::runtime_error(vlo_last_error(imu_handle()));
        // what the CPU path consumed: samples with t <= startTime (IMUManager.cpp:35-40) and those with t < endTime (:46-54)
        while (!_buffer.empty() && _buffer.front().time < endTime) _buffer.pop_front();
        return CombinedImuFactor(
                X(currentIndex - 1), V(currentIndex - 1),
                X(currentIndex), V(currentIndex),
                B(currentIndex - 1), B(currentIndex),
                Adopt(_integrator, bias, f)
                );
    }

    CombinedImuFactor IMUManager::getFactor(double endTime, uint64_t currentIndex, imuBias::ConstantBias bias)
    {
        return getFactor(_buffer.front().time, endTime, currentIndex, bias);
    }
}
