// BlockUtilities.cpp -- see CentroidalMPCWalking/BlockUtilities.h
#include "CentroidalMPCWalking/BlockUtilities.h"

#include <algorithm>

namespace CentroidalMPCWalking {

using BipedalLocomotion::Contacts::ContactListMap;
using BipedalLocomotion::Contacts::ContactPhaseList;
using BipedalLocomotion::Contacts::PlannedContact;

bool updateContactPhaseList(const std::chrono::nanoseconds& currentTime, const ContactPhaseList& plannerPhaseList,
                            const ContactPhaseList& mpcPhaseList, ContactPhaseList& contactPhaseList)
{
    ContactListMap merged;
    for (const auto& [name, plannerList] : plannerPhaseList.lists()) {
        // every contact of the planner that has not started yet
        for (auto it = plannerList.getNextContact(currentTime); it != plannerList.cend(); ++it)
            if (!merged[name].addContact(*it)) return false;
        // the contact the foot stands on now: where the MPC put it, for as long as the planner says
        const auto mpcIt = mpcPhaseList.lists().find(name);
        if (mpcIt == mpcPhaseList.lists().end()) return false;
        const auto mpcPresent = mpcIt->second.getActiveContact(currentTime);
        if (mpcPresent == mpcIt->second.cend()) continue;  // the foot is in the air for the MPC: nothing to keep
        const auto plannerPresent = plannerList.getActiveContact(currentTime);
        if (plannerPresent == plannerList.cend()) return false;
        PlannedContact contact = *mpcPresent;
        contact.activationTime = plannerPresent->activationTime;
        contact.deactivationTime = plannerPresent->deactivationTime;
        if (!merged[name].addContact(contact)) return false;
    }
    return contactPhaseList.setLists(merged);
}

bool resampleLinear(const std::vector<std::chrono::nanoseconds>& inputTimes, const std::vector<Eigen::Vector3d>& points,
                    const std::vector<std::chrono::nanoseconds>& outputTimes, std::vector<Eigen::Vector3d>& output)
{
    if (inputTimes.size() != points.size() || inputTimes.empty()) return false;
    for (size_t i = 1; i < inputTimes.size(); ++i)
        if (inputTimes[i] <= inputTimes[i - 1]) return false;
    output.resize(outputTimes.size());
    size_t seg = 0;
    for (size_t o = 0; o < outputTimes.size(); ++o) {
        const auto t = outputTimes[o];
        if (o > 0 && t < outputTimes[o - 1]) return false;  // ordered points only
        if (t <= inputTimes.front()) { output[o] = points.front(); continue; }
        if (t >= inputTimes.back()) { output[o] = points.back(); continue; }
        while (seg + 1 < inputTimes.size() && inputTimes[seg + 1] < t) ++seg;
        const double span = std::chrono::duration<double>(inputTimes[seg + 1] - inputTimes[seg]).count();
        const double a = std::chrono::duration<double>(t - inputTimes[seg]).count() / span;
        output[o] = points[seg] * (1.0 - a) + points[seg + 1] * a;
    }
    return true;
}

bool computeDesiredZMP(const std::map<std::string, BipedalLocomotion::Contacts::DiscreteGeometryContact>& contacts, double* zmp,
                       double halfLength, double halfWidth)
{
    zmp[0] = zmp[1] = 0.0;
    if (contacts.empty()) return false;
    double totalZ = 0.0;
    for (const auto& [name, contact] : contacts) {
        Eigen::Vector3d force, torque;
        const Eigen::Matrix3d Rt = contact.pose.rotation().transpose();
        for (const auto& corner : contact.corners) {
            force += corner.force;
            torque += corner.position.cross(Rt * corner.force);  // moment about the sole origin, contact frame
        }
        if (force[2] <= 0.001) continue;
        Eigen::Vector3d local(-torque[1] / force[2], torque[0] / force[2], 0.0);
        local[0] = std::min(halfLength, std::max(-halfLength, local[0]));
        local[1] = std::min(halfWidth, std::max(-halfWidth, local[1]));
        const Eigen::Vector3d world = contact.pose.act(local);
        zmp[0] += force[2] * world[0];
        zmp[1] += force[2] * world[1];
        totalZ += force[2];
    }
    if (totalZ < 0.001) return false;
    zmp[0] /= totalZ;
    zmp[1] /= totalZ;
    return true;
}

}  // namespace CentroidalMPCWalking
