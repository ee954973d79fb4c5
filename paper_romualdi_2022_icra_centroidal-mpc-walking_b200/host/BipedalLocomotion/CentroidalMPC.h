// CentroidalMPC.h -- host-side operator of the B200 centroidal-MPC solve, a drop-in for
// BipedalLocomotion::ReducedModelControllers::CentroidalMPC as the reference uses it:
//   member      src/centroidal-mpc-walking/include/CentroidalMPCWalking/CentroidalMPCBlock.h:72
//   initialize  src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:144   (group CENTROIDAL_MPC of centroidal_mpc.ini)
//   setState    :407      setReferenceTrajectory :579      setContactPhaseList :609
//   advance     :615      getOutput :598, :622, :626
// Same names, argument meaning and error behaviour (every call returns bool, logs, never throws; calling anything before
// initialize() fails).  advance() fills the solver's formal input (p, lbg, ubg, warm start) exactly as documented in
// SURVEY.md 8(a) a-7 and hands it to libcmpc_b200.so through the C ABI (include/cmpc_b200.h) -- there is no CPU solve.
// Added over the reference API: advanceBatch() (N independent controllers in ONE GPU launch) and solver statistics.
#pragma once

#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "BipedalLocomotion/Contacts.h"
#include "BipedalLocomotion/Math.h"
#include "BipedalLocomotion/ParametersHandler.h"

struct cmpc_config;

namespace BipedalLocomotion {
namespace ReducedModelControllers {

struct CentroidalMPCOutput {
    std::map<std::string, Contacts::DiscreteGeometryContact> contacts;       // knot-0 pose and corner forces
    std::map<std::string, Contacts::PlannedContact> nextPlannedContact;      // adjusted landing pose of the next activation
    Contacts::ContactPhaseList contactPhaseList;                             // input list with the adjusted contact edited in
    std::vector<Eigen::Vector3d> comTrajectory;                              // N + 1 knots
    std::vector<Eigen::Vector3d> comVelocityTrajectory;
    std::vector<Eigen::Vector3d> angularMomentumTrajectory;
    std::chrono::nanoseconds computationalTime{0};
};

struct CentroidalMPCSolverStats {
    int status{-1};        // CMPC_STATUS_* of include/cmpc_b200.h
    int iterations{0};
    double objective{0.0};
};

class CentroidalMPC {
public:
    CentroidalMPC();
    ~CentroidalMPC();
    CentroidalMPC(const CentroidalMPC&) = delete;
    CentroidalMPC& operator=(const CentroidalMPC&) = delete;

    bool initialize(std::weak_ptr<const ParametersHandler::IParametersHandler> handler);
    bool setContactPhaseList(const Contacts::ContactPhaseList& contactPhaseList);
    bool setState(const Eigen::Vector3d& com, const Eigen::Vector3d& dcom, const Eigen::Vector3d& angularMomentum);
    bool setState(const Eigen::Vector3d& com, const Eigen::Vector3d& dcom, const Eigen::Vector3d& angularMomentum,
                  const Math::Wrenchd& externalWrench);
    bool setReferenceTrajectory(const std::vector<Eigen::Vector3d>& com, const std::vector<Eigen::Vector3d>& angularMomentum);
    bool advance();
    const CentroidalMPCOutput& getOutput() const;
    bool isOutputValid() const;

    // ---- additions of the B200 build
    // one GPU launch for all controllers (they must have been initialised from the same configuration)
    static bool advanceBatch(const std::vector<CentroidalMPC*>& controllers);
    // the solver's formal input for the current tick in the reference's CasADi order: p[50N+27],
    // lbg/ubg[53N+15], x0[45N+15].  Host-only (no GPU needed): what advance() would hand to the solver.
    bool getSolverInputs(std::vector<double>& p, std::vector<double>& lbg, std::vector<double>& ubg, std::vector<double>& x0) const;
    // the compact tick record advance() uploads for the current tick (cmpc_tick_stride(horizon) doubles, include/cmpc_b200.h):
    // state, wrench, references and the contact windows; the device expands it into the formal input.  Host-only.
    bool getTickRecord(double* tick) const;
    const CentroidalMPCSolverStats& getSolverStats() const;
    bool getConfig(cmpc_config& cfg) const;
    int horizon() const;
    std::chrono::nanoseconds samplingTime() const;
    std::chrono::nanoseconds currentTime() const;
    const std::string& lastError() const;

private:
    struct Impl;
    std::unique_ptr<Impl> m_pimpl;
};

}  // namespace ReducedModelControllers
}  // namespace BipedalLocomotion
