// BlockUtilities.h -- the callers and data adapters on either side of the MPC solve, restated
// on the stand-in types of this library:
//   updateContactPhaseList   src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:32-110: the planner's future contacts
//                            plus the MPC's own (already adjusted) current contact with the planner's timing
//   resampleLinear           the Math::LinearSpline frequency adapter of CentroidalMPCBlock.cpp:201-260, 544-577: planner
//                            samples (50 Hz x slow-down) -> the N + 1 knots of the MPC
//   computeDesiredZMP        src/centroidal-mpc-walking/src/WholeBodyQPBlock.cpp:805-873: zero-moment point from the corner
//                            forces of the MPC output (per-contact local ZMP clamped to the foot, force-weighted average)
#pragma once

#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "BipedalLocomotion/Contacts.h"
#include "BipedalLocomotion/Math.h"

namespace CentroidalMPCWalking {

bool updateContactPhaseList(const std::chrono::nanoseconds& currentTime,
                            const BipedalLocomotion::Contacts::ContactPhaseList& plannerPhaseList,
                            const BipedalLocomotion::Contacts::ContactPhaseList& mpcPhaseList,
                            BipedalLocomotion::Contacts::ContactPhaseList& contactPhaseList);

// piecewise-linear interpolation of `points` given at `inputTimes` (increasing) evaluated at `outputTimes` (ordered);
// outside the input range the end points are held
bool resampleLinear(const std::vector<std::chrono::nanoseconds>& inputTimes, const std::vector<Eigen::Vector3d>& points,
                    const std::vector<std::chrono::nanoseconds>& outputTimes, std::vector<Eigen::Vector3d>& output);

// zmp[2] in the inertial frame; halfLength / halfWidth = the clamp of the local ZMP (0.08 / 0.03 in the reference)
bool computeDesiredZMP(const std::map<std::string, BipedalLocomotion::Contacts::DiscreteGeometryContact>& contacts, double* zmp,
                       double halfLength = 0.08, double halfWidth = 0.03);

}  // namespace CentroidalMPCWalking
