// capi.cpp -- flat C entry points over the C++ host operator (BipedalLocomotion::ReducedModelControllers::CentroidalMPC),
// so that tests and non-C++ hosts (ctypes) can drive the same object the reference's CentroidalMPCBlock would own.
// Times are in seconds at this boundary and converted to std::chrono::nanoseconds.
#include <cmath>
#include <cstring>
#include <memory>
#include <sstream>

#include "BipedalLocomotion/CentroidalMPC.h"
#include "CentroidalMPCWalking/BlockUtilities.h"
#include "cmpc_b200.h"

using namespace BipedalLocomotion;
using ReducedModelControllers::CentroidalMPC;

namespace {
struct Host {
    std::shared_ptr<ParametersHandler::IniImplementation> ini;
    CentroidalMPC mpc;
    Contacts::ContactListMap pending;
    std::string error;
};
std::chrono::nanoseconds ns(double s) { return std::chrono::nanoseconds((long long)std::llround(s * 1e9)); }
}  // namespace

extern "C" {

// ini_path: a centroidal_mpc.ini or a top-level file; group_path: "/"-separated groups to descend, "" = top level
void* cmpch_create(const char* ini_path, const char* group_path)
{
    auto h = std::make_unique<Host>();
    h->ini = std::make_shared<ParametersHandler::IniImplementation>();
    if (!h->ini->setFromFile(ini_path)) return nullptr;
    std::shared_ptr<ParametersHandler::IParametersHandler> cur = h->ini;
    std::stringstream ss(group_path ? group_path : "");
    std::string g;
    while (std::getline(ss, g, '/')) {
        if (g.empty()) continue;
        cur = cur->getGroup(g).lock();
        if (!cur) return nullptr;
    }
    if (!h->mpc.initialize(cur)) return nullptr;
    return h.release();
}
void cmpch_destroy(void* v) { delete static_cast<Host*>(v); }
int cmpch_horizon(void* v) { return static_cast<Host*>(v)->mpc.horizon(); }
double cmpch_sampling_time(void* v) { return std::chrono::duration<double>(static_cast<Host*>(v)->mpc.samplingTime()).count(); }
double cmpch_current_time(void* v) { return std::chrono::duration<double>(static_cast<Host*>(v)->mpc.currentTime()).count(); }
int cmpch_config(void* v, cmpc_config* out) { return static_cast<Host*>(v)->mpc.getConfig(*out) ? 0 : -1; }
const char* cmpch_last_error(void* v) { return static_cast<Host*>(v)->mpc.lastError().c_str(); }

int cmpch_set_state(void* v, const double* com, const double* dcom, const double* h, const double* wrench6)
{
    Math::Wrenchd w;
    if (wrench6) { for (int a = 0; a < 3; ++a) { w.force()[a] = wrench6[a]; w.torque()[a] = wrench6[3 + a]; } }
    return static_cast<Host*>(v)->mpc.setState({com[0], com[1], com[2]}, {dcom[0], dcom[1], dcom[2]}, {h[0], h[1], h[2]}, w) ? 0 : -1;
}
int cmpch_set_reference(void* v, int n, const double* com, const double* h)
{
    std::vector<Eigen::Vector3d> c(n), a(n);
    for (int i = 0; i < n; ++i) { c[i] = {com[3 * i], com[3 * i + 1], com[3 * i + 2]}; a[i] = {h[3 * i], h[3 * i + 1], h[3 * i + 2]}; }
    return static_cast<Host*>(v)->mpc.setReferenceTrajectory(c, a) ? 0 : -1;
}
// contact list of one foot: n contacts, activation / deactivation times [s], positions (3n), yaw angles (n)
int cmpch_set_contact_list(void* v, const char* name, int n, const double* t_on, const double* t_off, const double* pos, const double* yaw)
{
    Host* h = static_cast<Host*>(v);
    Contacts::ContactList list;
    list.setDefaultName(name);
    for (int i = 0; i < n; ++i)
        if (!list.addContact(manif::SE3d::fromYaw({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, yaw ? yaw[i] : 0.0), ns(t_on[i]), ns(t_off[i])))
            return -1;
    h->pending[name] = list;
    return 0;
}
int cmpch_commit_contacts(void* v, double force_sample_time)
{
    Host* h = static_cast<Host*>(v);
    Contacts::ContactPhaseList pl;
    pl.setLists(h->pending);
    if (force_sample_time > 0 && !pl.forceSampleTime(ns(force_sample_time))) return -1;
    return h->mpc.setContactPhaseList(pl) ? 0 : -1;
}
int cmpch_get_inputs(void* v, double* p, double* lbg, double* ubg, double* x0)
{
    std::vector<double> P, LB, UB, X;
    if (!static_cast<Host*>(v)->mpc.getSolverInputs(P, LB, UB, X)) return -1;
    std::memcpy(p, P.data(), 8 * P.size()); std::memcpy(lbg, LB.data(), 8 * LB.size());
    std::memcpy(ubg, UB.data(), 8 * UB.size()); std::memcpy(x0, X.data(), 8 * X.size());
    return 0;
}
// the compact tick record advance() would upload (cmpc_tick_stride(horizon) doubles)
int cmpch_get_tick(void* v, double* tick) { return static_cast<Host*>(v)->mpc.getTickRecord(tick) ? 0 : -1; }
int cmpch_advance(void* v) { return static_cast<Host*>(v)->mpc.advance() ? 0 : -1; }
int cmpch_advance_batch(void** hosts, int n)
{
    std::vector<CentroidalMPC*> c(n);
    for (int i = 0; i < n; ++i) c[i] = &static_cast<Host*>(hosts[i])->mpc;
    return CentroidalMPC::advanceBatch(c) ? 0 : -1;
}
int cmpch_is_output_valid(void* v) { return static_cast<Host*>(v)->mpc.isOutputValid() ? 1 : 0; }
// knot-0 output of one contact: position (3), rotation (9, column major), corner forces (12)
int cmpch_get_contact_output(void* v, const char* name, double* pos, double* rot, double* forces)
{
    const auto& out = static_cast<Host*>(v)->mpc.getOutput();
    auto it = out.contacts.find(name);
    if (it == out.contacts.end()) return -1;
    std::memcpy(pos, it->second.pose.translation().data(), 24);
    std::memcpy(rot, it->second.pose.rotation().data(), 72);
    for (size_t j = 0; j < it->second.corners.size(); ++j) std::memcpy(forces + 3 * j, it->second.corners[j].force.data(), 24);
    return 0;
}
// adjusted landing of the next activation: position (3), activation time [s]; 1 when there is none inside the horizon
int cmpch_get_next_planned_contact(void* v, const char* name, double* pos, double* t_on)
{
    const auto& out = static_cast<Host*>(v)->mpc.getOutput();
    auto it = out.nextPlannedContact.find(name);
    if (it == out.nextPlannedContact.end()) return 1;
    std::memcpy(pos, it->second.pose.translation().data(), 24);
    *t_on = std::chrono::duration<double>(it->second.activationTime).count();
    return 0;
}
int cmpch_get_trajectories(void* v, double* com, double* dcom, double* h)
{
    const auto& out = static_cast<Host*>(v)->mpc.getOutput();
    for (size_t k = 0; k < out.comTrajectory.size(); ++k) {
        std::memcpy(com + 3 * k, out.comTrajectory[k].data(), 24);
        std::memcpy(dcom + 3 * k, out.comVelocityTrajectory[k].data(), 24);
        std::memcpy(h + 3 * k, out.angularMomentumTrajectory[k].data(), 24);
    }
    return (int)out.comTrajectory.size();
}
int cmpch_get_stats(void* v, int* status, int* iterations, double* objective)
{
    const auto& s = static_cast<Host*>(v)->mpc.getSolverStats();
    *status = s.status; *iterations = s.iterations; *objective = s.objective;
    return 0;
}
// contact list of the output phase list (the adjusted contact edited in): up to `cap` contacts
int cmpch_get_output_contact_list(void* v, const char* name, int cap, double* t_on, double* t_off, double* pos)
{
    const auto& lists = static_cast<Host*>(v)->mpc.getOutput().contactPhaseList.lists();
    auto it = lists.find(name);
    if (it == lists.end()) return -1;
    int n = 0;
    for (const auto& c : it->second) {
        if (n >= cap) break;
        t_on[n] = std::chrono::duration<double>(c.activationTime).count();
        t_off[n] = std::chrono::duration<double>(c.deactivationTime).count();
        std::memcpy(pos + 3 * n, c.pose.translation().data(), 24);
        ++n;
    }
    return n;
}

// ZMP of the current output (CentroidalMPCWalking::computeDesiredZMP)
int cmpch_desired_zmp(void* v, double* zmp2)
{
    return CentroidalMPCWalking::computeDesiredZMP(static_cast<Host*>(v)->mpc.getOutput().contacts, zmp2) ? 0 : -1;
}
// planner list (pending lists set with cmpch_set_contact_list) merged with the MPC's output list at the controller's
// current time, then handed to setContactPhaseList: the per-tick sequence of CentroidalMPCBlock::advance()
int cmpch_commit_contacts_merged(void* v, double force_sample_time, int first_run)
{
    Host* h = static_cast<Host*>(v);
    Contacts::ContactPhaseList planner, merged;
    planner.setLists(h->pending);
    if (force_sample_time > 0 && !planner.forceSampleTime(ns(force_sample_time))) return -1;
    if (first_run) merged = planner;
    else if (!CentroidalMPCWalking::updateContactPhaseList(h->mpc.currentTime(), planner, h->mpc.getOutput().contactPhaseList, merged)) return -2;
    return h->mpc.setContactPhaseList(merged) ? 0 : -1;
}
// Math::LinearSpline-style resampling: n_in samples (times [s], points 3 n_in) -> n_out points
int cmpch_resample_linear(int n_in, const double* t_in, const double* p_in, int n_out, const double* t_out, double* p_out)
{
    std::vector<std::chrono::nanoseconds> ti(n_in), to(n_out);
    std::vector<Eigen::Vector3d> pi(n_in), po;
    for (int i = 0; i < n_in; ++i) { ti[i] = ns(t_in[i]); pi[i] = {p_in[3 * i], p_in[3 * i + 1], p_in[3 * i + 2]}; }
    for (int i = 0; i < n_out; ++i) to[i] = ns(t_out[i]);
    if (!CentroidalMPCWalking::resampleLinear(ti, pi, to, po)) return -1;
    for (int i = 0; i < n_out; ++i) std::memcpy(p_out + 3 * i, po[i].data(), 24);
    return 0;
}

}  // extern "C"
