// TEST DRIVER for the link-time IMU seam (tests/test_abi_and_host.py): a caller written like GraphManager::reserveNode
// (GraphManager.cpp:51-69) -- it sees only the reference's IMUManager.h and calls getFactor through a plain pointer --
// linked with IMUManager_vlo.o instead of IMUManager.o.  With a GPU it also prints the preintegrated velocity / position
// of the reference's known-answer case (gtsam_fusion/test/UnitTests.cpp:30-66: dv = 0.0175, dp = 0.0011875).
#include <gtsam_fusion/IMUManager.h>
#include <cstdio>
int main(int argc, char **)
{
    if (argc < 2) return 0;                    // link check only (no GPU needed)
    auto prm = boost::shared_ptr<gtsam::PreintegratedCombinedMeasurements::Params>(new gtsam::PreintegratedCombinedMeasurements::Params());
    VILFusion::IMUManager mgr(prm);
    for (int k = 0; k <= 2; k++) {             // samples (t, a = w) = (0, 0), (0.1, 0.1), (0.2, 0.2); window [0, 0.15]
        gtsam::Vector3 v;
        v[0] = v[1] = v[2] = 0.1 * k;
        mgr.addIMUMeasurement(0.1 * k, v, v);
    }
    gtsam::CombinedImuFactor f = mgr.getFactor(0.0, 0.15, 1, gtsam::imuBias::ConstantBias());
    const auto &pim = f.preintegratedMeasurements();
    std::printf("%.10f %.10f %.10f\n", pim.deltaTij(), pim.deltaXij().v_[0], pim.deltaXij().t_[0]);
    return 0;
}
