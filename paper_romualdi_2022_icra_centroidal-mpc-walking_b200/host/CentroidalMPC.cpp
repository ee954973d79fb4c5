// CentroidalMPC.cpp -- see BipedalLocomotion/CentroidalMPC.h.
//
// What lives here is the host side of the path: configuration (both ini dialects of the reference), the
// population of the solver's formal input from (state, references, contact phase list), the warm-start
// bookkeeping and the unpacking of the solution (a-9).  The solve itself is cmpc_solve_host() of libcmpc_b200.so.
#include "BipedalLocomotion/CentroidalMPC.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "cmpc_b200.h"

namespace BipedalLocomotion {
namespace ReducedModelControllers {

using namespace std::chrono_literals;
using Contacts::ContactList;
using Contacts::PlannedContact;

namespace {
constexpr double kInf = 1e20;        // |bound| >= 1e19 is "no bound" for the solver (IPOPT nlp_upper_bound_inf)
constexpr double kGravity = 9.80665;  // tmp.c:3916
constexpr int NC = CMPC_NUM_CONTACTS, NJ = CMPC_NUM_CORNERS;

// index layout of x / p / g in the reference's CasADi order (tmp.c:62-67)
struct Layout {
    int N;
    int n() const { return 45 * N + 15; }
    int np() const { return 50 * N + 27; }
    int m() const { return 53 * N + 15; }
    int x_com(int k) const { return 3 * k; }
    int x_dcom(int k) const { return 3 * (N + 1) + 3 * k; }
    int x_h(int k) const { return 6 * (N + 1) + 3 * k; }
    int x_cbase(int c) const { return 9 * (N + 1) + c * (18 * N + 3); }
    int x_pos(int c, int k) const { return x_cbase(c) + 3 * k; }
    int x_vel(int c, int k) const { return x_cbase(c) + 3 * (N + 1) + 3 * k; }
    int x_frc(int c, int j, int k) const { return x_cbase(c) + 6 * N + 3 + 3 * N * j + 3 * k; }
    int p_cbase(int c) const { return c * (19 * N + 6); }
    int p_rot(int c, int k) const { return p_cbase(c) + 9 * k; }
    int p_upper(int c, int k) const { return p_cbase(c) + 9 * N + 3 * k; }
    int p_lower(int c, int k) const { return p_cbase(c) + 12 * N + 3 * k; }
    int p_en(int c, int k) const { return p_cbase(c) + 15 * N + k; }
    int p_nom(int c, int k) const { return p_cbase(c) + 16 * N + 3 * k; }
    int p_cur(int c) const { return p_cbase(c) + 19 * N + 3; }
    int p_glob() const { return 38 * N + 12; }
    int p_comref(int k) const { return p_glob() + 9 + 3 * k; }
    int p_href(int k) const { return p_glob() + 9 + 3 * (N + 1) + 3 * k; }
    int p_extf(int k) const { return p_glob() + 9 + 6 * (N + 1) + 3 * k; }
    int p_extt(int k) const { return p_glob() + 9 + 6 * (N + 1) + 3 * N + 3 * k; }
    int g_com(int k) const { return 15 + 3 * k; }
    int g_dcom(int k) const { return 15 + 3 * N + 3 * k; }
    int g_h(int k) const { return 15 + 6 * N + 3 * k; }
    int g_pos(int c, int k) const { return 15 + 9 * N + 3 * N * c + 3 * k; }
    int g_box(int c, int k) const { return 15 + 15 * N + c * 19 * N + 3 * k; }
    int g_fric(int c, int j, int k) const { return 15 + 15 * N + c * 19 * N + 3 * N + 16 * k + 4 * j; }
};

struct ContactConfig {
    std::string name;
    double corners[NJ][3];
    double boxUpper[3], boxLower[3];
};

std::chrono::nanoseconds toNs(double seconds) { return std::chrono::nanoseconds((long long)std::llround(seconds * 1e9)); }
}  // namespace

enum class State { NotInitialized, Initialized, OutputValid, OutputInvalid };

struct CentroidalMPC::Impl {
    State fsm{State::NotInitialized};
    cmpc_config cfg{};
    Layout L{12};
    std::chrono::nanoseconds dT{100ms}, currentTime{0};
    bool warmStartEnabled{false}, stepAdjustmentEnabled{true};
    int verbosity{0};
    std::string linearSolver{"mumps"}, solverName{"ipopt"};
    std::vector<ContactConfig> contacts;  // sorted by name: index c of the NLP (std::map order of the reference)

    // inputs of the current tick
    bool hasState{false}, hasReference{false}, hasContacts{false};
    Eigen::Vector3d com, dcom, angMom;
    Math::Wrenchd wrench;
    std::vector<Eigen::Vector3d> comRef, angMomRef;
    Contacts::ContactPhaseList phaseList;

    // solver
    cmpc_handle handle{nullptr};
    bool hasPrevious{false};
    std::vector<double> xPrev, lamPrev;  // solution of the previous tick (CasADi order)
    std::vector<double> p, lbg, ubg, x, lam;
    CentroidalMPCOutput output;
    CentroidalMPCSolverStats stats;
    std::string error;

    bool fail(const std::string& what)
    {
        error = what;
        if (verbosity >= 0) std::fprintf(stderr, "[CentroidalMPC] %s\n", what.c_str());
        return false;
    }
    int contactIndex(const std::string& name) const
    {
        for (size_t c = 0; c < contacts.size(); ++c)
            if (contacts[c].name == name) return (int)c;
        return -1;
    }
    bool loadParameters(const ParametersHandler::IParametersHandler& h);
    bool fillInputs(std::vector<double>& p_, std::vector<double>& lbg_, std::vector<double>& ubg_, std::vector<double>& x0_,
                    std::vector<double>* lam0_) const;
    // the compact tick record of include/cmpc_b200.h (what advance() uploads; the device expands it: cmpc_populate)
    bool fillTick(double* tick);
    // knot-wise contact activity and the knot-0 rotations of the tick in flight (unpack needs them; p stays on the device)
    std::vector<char> enabled[NC];
    Eigen::Matrix3d rot0[NC];
    // controllers whose previous solutions are resident on the device of THIS controller's handle, in batch order
    std::vector<const CentroidalMPC*> residentBatch;
    bool ensureHandle();
    void unpack(std::chrono::nanoseconds elapsed);
};

// ---------------------------------------------------------------------------------------------------- configuration
bool CentroidalMPC::Impl::loadParameters(const ParametersHandler::IParametersHandler& h)
{
    cmpc_default_config(&cfg);
    double dt = 0.0;
    // dialect "2023" (ergoCub*): sampling_time + time_horizon in seconds; dialect "2022" (iCub*): controller_* keys
    if (h.getParameter("sampling_time", dt)) {
        double horizonTime = 0.0;
        if (!h.getParameter("time_horizon", horizonTime)) return fail("parameter time_horizon not found");
        dT = toNs(dt);
        if (dT <= 0ns) return fail("sampling_time must be positive");
        cfg.horizon = (int)(toNs(horizonTime) / dT);  // integer division of nanoseconds, as BLF does
    } else if (h.getParameter("controller_sampling_time", dt)) {
        int horizon = 0;
        if (!h.getParameter("controller_horizon", horizon)) return fail("parameter controller_horizon not found");
        dT = toNs(dt);
        cfg.horizon = horizon;
    } else {
        return fail("parameter sampling_time (or controller_sampling_time) not found");
    }
    if (cfg.horizon < 2) return fail("the horizon must contain at least 2 knots");
    cfg.sampling_time = std::chrono::duration<double>(dT).count();
    L.N = cfg.horizon;

    int maxContacts = 0;
    if (!h.getParameter("number_of_maximum_contacts", maxContacts)) return fail("parameter number_of_maximum_contacts not found");
    if (maxContacts != NC) return fail("number_of_maximum_contacts must be 2 (left_foot, right_foot)");
    if (!h.getParameter("number_of_slices", cfg.number_of_slices)) return fail("parameter number_of_slices not found");
    if (cfg.number_of_slices != 1) return fail("only number_of_slices 1 is supported");
    if (!h.getParameter("static_friction_coefficient", cfg.static_friction_coefficient))
        return fail("parameter static_friction_coefficient not found");

    std::vector<double> v;
    if (!h.getParameter("com_weight", v) || v.size() != 3) return fail("parameter com_weight (3) not found");
    std::copy(v.begin(), v.end(), cfg.com_weight);
    if (!h.getParameter("contact_position_weight", cfg.contact_position_weight)) return fail("parameter contact_position_weight not found");
    if (!h.getParameter("force_rate_of_change_weight", v) || v.size() != 3) return fail("parameter force_rate_of_change_weight (3) not found");
    std::copy(v.begin(), v.end(), cfg.force_rate_of_change_weight);
    if (!h.getParameter("angular_momentum_weight", cfg.angular_momentum_weight)) return fail("parameter angular_momentum_weight not found");
    cfg.contact_force_symmetry_weight = 0.0;  // the 2022 NLP had no symmetry term
    h.getParameter("contact_force_symmetry_weight", cfg.contact_force_symmetry_weight);

    // solver options
    h.getParameter("linear_solver", linearSolver);  // mumps / ma97 / ma27: accepted, the KKT system is solved by the Riccati kernel
    h.getParameter("solver_name", solverName);
    if (solverName != "ipopt") return fail("solver_name \"" + solverName + "\" is not supported: only the ipopt path is restated");
    cfg.ipopt_tolerance = 1e-8;
    h.getParameter("ipopt_tolerance", cfg.ipopt_tolerance);
    int maxIter = 0;
    if (h.getParameter("ipopt_max_iteration", maxIter) && maxIter > 0) cfg.ipopt_max_iteration = maxIter;
    h.getParameter("nlp_scaling_max_gradient", cfg.nlp_scaling_max_gradient);  // additions: IPOPT option names, IPOPT defaults
    h.getParameter("acceptable_tol", cfg.acceptable_tol);
    h.getParameter("acceptable_iter", cfg.acceptable_iter);
    h.getParameter("solver_verbosity", verbosity);
    h.getParameter("is_warm_start_enabled", warmStartEnabled);
    h.getParameter("step_adjustment_enabled", stepAdjustmentEnabled);  // addition: false = every step box has zero width
    h.getParameter("cuda_device", cfg.device);                         // addition
    {   // addition: barrier update of the batched interior point ("mehrotra" = library default, "monotone" = IPOPT's own path)
        std::string muStrategy;
        if (h.getParameter("mu_strategy", muStrategy)) {
            if (muStrategy == "monotone") cfg.mu_strategy = CMPC_MU_MONOTONE;
            else if (muStrategy == "mehrotra") cfg.mu_strategy = CMPC_MU_MEHROTRA;
            else return fail("mu_strategy must be \"monotone\" or \"mehrotra\"");
        }
    }

    contacts.clear();
    for (int i = 0; i < maxContacts; ++i) {
        auto gw = h.getGroup("CONTACT_" + std::to_string(i));
        auto g = gw.lock();
        if (!g) return fail("group CONTACT_" + std::to_string(i) + " not found");
        ContactConfig cc;
        if (!g->getParameter("contact_name", cc.name)) return fail("contact_name not found in CONTACT_" + std::to_string(i));
        int corners = 0;
        if (!g->getParameter("number_of_corners", corners) || corners != NJ) return fail("number_of_corners must be 4");
        for (int j = 0; j < NJ; ++j) {
            if (!g->getParameter("corner_" + std::to_string(j), v) || v.size() != 3)
                return fail("corner_" + std::to_string(j) + " (3) not found in CONTACT_" + std::to_string(i));
            std::copy(v.begin(), v.end(), cc.corners[j]);
        }
        if (!g->getParameter("bounding_box_upper_limit", v) || v.size() != 3) return fail("bounding_box_upper_limit (3) not found");
        std::copy(v.begin(), v.end(), cc.boxUpper);
        if (!g->getParameter("bounding_box_lower_limit", v) || v.size() != 3) return fail("bounding_box_lower_limit (3) not found");
        std::copy(v.begin(), v.end(), cc.boxLower);
        contacts.push_back(cc);
    }
    std::sort(contacts.begin(), contacts.end(), [](const ContactConfig& a, const ContactConfig& b) { return a.name < b.name; });
    if (contacts[0].name == contacts[1].name) return fail("the two contacts must have different names");
    for (int c = 0; c < NC; ++c) {
        std::memcpy(cfg.corners[c], contacts[c].corners, sizeof(double) * NJ * 3);
        std::memcpy(cfg.bounding_box_upper_limit[c], contacts[c].boxUpper, sizeof(double) * 3);
        std::memcpy(cfg.bounding_box_lower_limit[c], contacts[c].boxLower, sizeof(double) * 3);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------- formal input
bool CentroidalMPC::Impl::fillInputs(std::vector<double>& p_, std::vector<double>& lbg_, std::vector<double>& ubg_,
                                     std::vector<double>& x0_, std::vector<double>* lam0_) const
{
    const int N = L.N;
    p_.assign(L.np(), 0.0);
    lbg_.assign(L.m(), 0.0);
    ubg_.assign(L.m(), 0.0);
    x0_.assign(L.n(), 0.0);
    const auto t0 = currentTime;

    for (int c = 0; c < NC; ++c) {
        const auto lit = phaseList.lists().find(contacts[c].name);
        if (lit == phaseList.lists().end()) return false;
        const ContactList& list = lit->second;
        std::vector<ContactList::const_iterator> act(N + 1);
        for (int k = 0; k <= N; ++k) act[k] = list.getActiveContact(t0 + k * dT);
        const auto end = list.cend();
        const auto before = list.getActiveContact(t0 - dT);  // contact of the knot before the horizon

        std::vector<Eigen::Vector3d> nom(N + 1);
        for (int k = 0; k <= N; ++k) {
            Eigen::Matrix3d R = Eigen::Matrix3d::Identity();  // identity on swing knots
            if (act[k] != end) {
                nom[k] = act[k]->pose.translation();
                R = act[k]->pose.rotation();
            } else {
                const auto prev = k >= 1 ? act[k - 1] : before;
                if (prev != end) nom[k] = prev->pose.translation();  // first swing knot: the foot has not moved yet
                else {
                    const auto next = list.getNextContact(t0 + k * dT);
                    if (next != end) nom[k] = next->pose.translation();  // in the air: the landing position
                    else {
                        const auto last = list.getPresentContact(t0 + k * dT);
                        if (last != end) nom[k] = last->pose.translation();
                    }
                }
            }
            std::memcpy(&p_[L.p_nom(c, k)], nom[k].data(), 24);
            std::memcpy(&x0_[L.x_pos(c, k)], nom[k].data(), 24);
            if (k < N) {
                std::memcpy(&p_[L.p_rot(c, k)], R.data(), 72);  // vec(R), column major
                p_[L.p_en(c, k)] = act[k] != end ? 1.0 : 0.0;
            }
        }
        // current position of the foot
        Eigen::Vector3d cur = nom[0];
        if (act[0] == end && before == end) {
            const auto prev = list.getPresentContact(t0);
            const auto next = list.getNextContact(t0);
            if (prev != end && next != end) {  // mid swing: between lift-off and landing
                const double span = std::chrono::duration<double>(next->activationTime - prev->deactivationTime).count();
                const double prog = span > 0 ? std::chrono::duration<double>(t0 - prev->deactivationTime).count() / span : 1.0;
                cur = prev->pose.translation() + (next->pose.translation() - prev->pose.translation()) * prog;
            }
        }
        std::memcpy(&p_[L.p_cur(c)], cur.data(), 24);
        for (int a = 0; a < 3; ++a) lbg_[9 + 3 * c + a] = ubg_[9 + 3 * c + a] = cur[a];
        // step-adjustment box rows of knot k: R_k' (pos_{k+1} - nominal_{k+1}) in [lower_k, upper_k]
        for (int k = 0; k < N; ++k) {
            double lo[3] = {0, 0, 0}, up[3] = {0, 0, 0};
            if (stepAdjustmentEnabled) {
                if (act[k] == end) {  // swing knot: the foot is free
                    for (int a = 0; a < 3; ++a) { lo[a] = -kInf; up[a] = kInf; }
                } else if (!(act[0] != end && act[k] == act[0])) {  // a contact of the future: adjustable inside its box
                    for (int a = 0; a < 3; ++a) { lo[a] = contacts[c].boxLower[a]; up[a] = contacts[c].boxUpper[a]; }
                }  // the contact the foot stands on now cannot move: zero width
            }
            for (int a = 0; a < 3; ++a) {
                p_[L.p_upper(c, k) + a] = up[a]; p_[L.p_lower(c, k) + a] = lo[a];
                lbg_[L.g_box(c, k) + a] = lo[a]; ubg_[L.g_box(c, k) + a] = up[a];
            }
            for (int j = 0; j < NJ; ++j)
                for (int r = 0; r < 4; ++r) lbg_[L.g_fric(c, j, k) + r] = -kInf;  // A R' f <= 0
            for (int j = 0; j < NJ; ++j) x0_[L.x_frc(c, j, k) + 2] = kGravity / (NC * NJ);  // cold start: weight shared by the corners
        }
    }
    // state, references, external wrench (column 0 only)
    const int g0 = L.p_glob();
    for (int a = 0; a < 3; ++a) {
        p_[g0 + a] = com[a]; p_[g0 + 3 + a] = dcom[a]; p_[g0 + 6 + a] = angMom[a];
        lbg_[a] = ubg_[a] = com[a]; lbg_[3 + a] = ubg_[3 + a] = dcom[a]; lbg_[6 + a] = ubg_[6 + a] = angMom[a];
        p_[L.p_extf(0) + a] = wrench.force()[a];
        p_[L.p_extt(0) + a] = wrench.torque()[a];
    }
    for (int k = 0; k <= N; ++k) {
        std::memcpy(&p_[L.p_comref(k)], comRef[k].data(), 24);
        std::memcpy(&p_[L.p_href(k)], angMomRef[k].data(), 24);
        std::memcpy(&x0_[L.x_com(k)], comRef[k].data(), 24);
    }
    // warm start: the previous solution moved one knot towards the present, last knot repeated
    if (lam0_) lam0_->assign(L.m(), 0.0);
    if (warmStartEnabled && hasPrevious) {
        auto shift3 = [&](std::vector<double>& dst, const std::vector<double>& src, int base, int cols) {
            for (int k = 0; k < cols; ++k)
                for (int a = 0; a < 3; ++a) dst[base + 3 * k + a] = src[base + 3 * std::min(k + 1, cols - 1) + a];
        };
        shift3(x0_, xPrev, L.x_com(0), N + 1); shift3(x0_, xPrev, L.x_dcom(0), N + 1); shift3(x0_, xPrev, L.x_h(0), N + 1);
        for (int c = 0; c < NC; ++c) {
            shift3(x0_, xPrev, L.x_pos(c, 0), N + 1);
            shift3(x0_, xPrev, L.x_vel(c, 0), N);
            for (int j = 0; j < NJ; ++j) shift3(x0_, xPrev, L.x_frc(c, j, 0), N);
        }
        if (lam0_) {
            shift3(*lam0_, lamPrev, L.g_com(0), N); shift3(*lam0_, lamPrev, L.g_dcom(0), N); shift3(*lam0_, lamPrev, L.g_h(0), N);
            for (int c = 0; c < NC; ++c) {
                shift3(*lam0_, lamPrev, L.g_pos(c, 0), N);
                shift3(*lam0_, lamPrev, L.g_box(c, 0), N);
                for (int k = 0; k < N; ++k)
                    for (int r = 0; r < 16; ++r) (*lam0_)[L.g_fric(c, 0, k) + r] = lamPrev[L.g_fric(c, 0, std::min(k + 1, N - 1)) + r];
            }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------- tick record
bool CentroidalMPC::Impl::fillTick(double* tick)
{
    const int N = L.N;
    const int stride = cmpc_tick_stride(N);
    std::fill(tick, tick + stride, 0.0);
    for (int a = 0; a < 3; ++a) {
        tick[a] = com[a]; tick[3 + a] = dcom[a]; tick[6 + a] = angMom[a];
        tick[9 + a] = wrench.force()[a]; tick[12 + a] = wrench.torque()[a];
    }
    tick[15] = stepAdjustmentEnabled ? 1.0 : 0.0;
    for (int k = 0; k <= N; ++k) {
        std::memcpy(tick + 17 + 3 * k, comRef[k].data(), 24);
        std::memcpy(tick + 17 + 3 * (N + 1) + 3 * k, angMomRef[k].data(), 24);
    }
    const auto t0 = currentTime;
    const auto horizonEnd = t0 + N * dT;
    constexpr int kMaxContacts = 6, kRecord = 14;
    for (int c = 0; c < NC; ++c) {
        const auto lit = phaseList.lists().find(contacts[c].name);
        if (lit == phaseList.lists().end()) return false;
        const ContactList& list = lit->second;
        // window: from the contact the foot stands (or last stood) on -- one earlier if that one was still active a sampling
        // time ago -- up to the first contact that starts after the horizon
        auto start = list.getPresentContact(t0);
        if (start == list.cend()) start = list.cbegin();
        else if (start != list.cbegin()) {
            auto prev = start; --prev;
            if (prev->isContactActive(t0 - dT)) start = prev;
        }
        double* out = tick + 17 + 6 * (N + 1) + c * (1 + kMaxContacts * kRecord);
        int n = 0;
        for (auto it = start; it != list.cend(); ++it) {
            if (n == kMaxContacts) return fail("more than 6 contacts of " + contacts[c].name + " inside the horizon");
            double* r = out + 1 + n * kRecord;
            const auto relOn = it->activationTime - t0;
            r[0] = (double)relOn.count();
            r[1] = it->deactivationTime == std::chrono::nanoseconds::max() ? 1e18 : (double)(it->deactivationTime - t0).count();
            std::memcpy(r + 2, it->pose.translation().data(), 24);
            std::memcpy(r + 5, it->pose.rotation().data(), 72);
            ++n;
            if (it->activationTime > horizonEnd) break;
        }
        out[0] = (double)n;
        // what unpack() needs of the formal input
        enabled[c].assign(N, 0);
        for (int k = 0; k < N; ++k) enabled[c][k] = list.getActiveContact(t0 + k * dT) != list.cend() ? 1 : 0;
        const auto a0 = list.getActiveContact(t0);
        rot0[c] = a0 != list.cend() ? a0->pose.rotation() : Eigen::Matrix3d::Identity();
    }
    return true;
}

bool CentroidalMPC::Impl::ensureHandle()
{
    if (handle) return true;
    const int rc = cmpc_create(&cfg, &handle);
    if (rc != CMPC_OK) return fail(std::string("cmpc_create failed: ") + cmpc_error_string(rc));
    return true;
}

// ---------------------------------------------------------------------------------------------------- output
void CentroidalMPC::Impl::unpack(std::chrono::nanoseconds elapsed)
{
    const int N = L.N;
    output.contacts.clear();
    output.nextPlannedContact.clear();
    output.comTrajectory.assign(N + 1, Eigen::Vector3d());
    output.comVelocityTrajectory.assign(N + 1, Eigen::Vector3d());
    output.angularMomentumTrajectory.assign(N + 1, Eigen::Vector3d());
    for (int k = 0; k <= N; ++k) {
        std::memcpy(output.comTrajectory[k].data(), &x[L.x_com(k)], 24);
        std::memcpy(output.comVelocityTrajectory[k].data(), &x[L.x_dcom(k)], 24);
        std::memcpy(output.angularMomentumTrajectory[k].data(), &x[L.x_h(k)], 24);
    }
    Contacts::ContactListMap lists = phaseList.lists();
    for (int c = 0; c < NC; ++c) {
        // knot 0: pose and corner forces (zero when the contact is not enabled)
        Contacts::DiscreteGeometryContact dc;
        dc.name = contacts[c].name;
        dc.index = c;
        Eigen::Vector3d pos;
        std::memcpy(pos.data(), &x[L.x_pos(c, 0)], 24);
        dc.pose = manif::SE3d(pos, rot0[c]);
        const double en0 = enabled[c][0] ? 1.0 : 0.0;
        dc.corners.resize(NJ);
        for (int j = 0; j < NJ; ++j) {
            dc.corners[j].position = Eigen::Vector3d(contacts[c].corners[j][0], contacts[c].corners[j][1], contacts[c].corners[j][2]);
            for (int a = 0; a < 3; ++a) dc.corners[j].force[a] = en0 * x[L.x_frc(c, j, 0) + a];
        }
        output.contacts[dc.name] = dc;
        // the next activation inside the horizon: its landing position is a decision variable of the MPC
        ContactList& list = lists[dc.name];
        int landing = -1;
        for (int k = 1; k < N; ++k)
            if (enabled[c][k] && !enabled[c][k - 1]) { landing = k; break; }
        if (landing > 0) {
            const auto it = list.getActiveContact(currentTime + landing * dT);
            if (it != list.cend()) {
                PlannedContact adj = *it;
                Eigen::Vector3d lp;
                // pos_{landing + 1} = pos_landing (the foot is in stance) is the row the step box constrains
                std::memcpy(lp.data(), &x[L.x_pos(c, std::min(landing + 1, N))], 24);
                adj.pose.translation(lp);
                output.nextPlannedContact[dc.name] = adj;
                list.editContact(it, adj);
            }
        }
    }
    output.contactPhaseList.setLists(lists);
    output.computationalTime = elapsed;
}

// ---------------------------------------------------------------------------------------------------- public API
CentroidalMPC::CentroidalMPC() : m_pimpl(std::make_unique<Impl>()) {}
CentroidalMPC::~CentroidalMPC()
{
    if (m_pimpl && m_pimpl->handle) cmpc_destroy(m_pimpl->handle);
}

bool CentroidalMPC::initialize(std::weak_ptr<const ParametersHandler::IParametersHandler> handler)
{
    auto ptr = handler.lock();
    if (!ptr) return m_pimpl->fail("initialize: the parameter handler is not valid");
    if (!m_pimpl->loadParameters(*ptr)) return false;
    m_pimpl->currentTime = 0ns;
    m_pimpl->hasPrevious = m_pimpl->hasState = m_pimpl->hasReference = m_pimpl->hasContacts = false;
    m_pimpl->wrench.setZero();
    m_pimpl->fsm = State::Initialized;
    return true;
}

bool CentroidalMPC::setContactPhaseList(const Contacts::ContactPhaseList& contactPhaseList)
{
    if (m_pimpl->fsm == State::NotInitialized) return m_pimpl->fail("setContactPhaseList: call initialize() first");
    if (contactPhaseList.size() == 0) return m_pimpl->fail("setContactPhaseList: the contact phase list is empty");
    for (const auto& c : m_pimpl->contacts)
        if (contactPhaseList.lists().find(c.name) == contactPhaseList.lists().end())
            return m_pimpl->fail("setContactPhaseList: no contact list for " + c.name);
    m_pimpl->phaseList = contactPhaseList;
    m_pimpl->hasContacts = true;
    return true;
}

bool CentroidalMPC::setState(const Eigen::Vector3d& com, const Eigen::Vector3d& dcom, const Eigen::Vector3d& angularMomentum)
{
    return setState(com, dcom, angularMomentum, Math::Wrenchd::Zero());
}

bool CentroidalMPC::setState(const Eigen::Vector3d& com, const Eigen::Vector3d& dcom, const Eigen::Vector3d& angularMomentum,
                             const Math::Wrenchd& externalWrench)
{
    if (m_pimpl->fsm == State::NotInitialized) return m_pimpl->fail("setState: call initialize() first");
    m_pimpl->com = com; m_pimpl->dcom = dcom; m_pimpl->angMom = angularMomentum; m_pimpl->wrench = externalWrench;
    m_pimpl->hasState = true;
    return true;
}

bool CentroidalMPC::setReferenceTrajectory(const std::vector<Eigen::Vector3d>& com, const std::vector<Eigen::Vector3d>& angularMomentum)
{
    if (m_pimpl->fsm == State::NotInitialized) return m_pimpl->fail("setReferenceTrajectory: call initialize() first");
    const size_t need = (size_t)m_pimpl->L.N + 1;
    if (com.size() < need || angularMomentum.size() < need)
        return m_pimpl->fail("setReferenceTrajectory: at least horizon + 1 = " + std::to_string(need) + " samples are required");
    m_pimpl->comRef.assign(com.begin(), com.begin() + need);
    m_pimpl->angMomRef.assign(angularMomentum.begin(), angularMomentum.begin() + need);
    m_pimpl->hasReference = true;
    return true;
}

bool CentroidalMPC::getSolverInputs(std::vector<double>& p, std::vector<double>& lbg, std::vector<double>& ubg, std::vector<double>& x0) const
{
    if (m_pimpl->fsm == State::NotInitialized || !m_pimpl->hasState || !m_pimpl->hasReference || !m_pimpl->hasContacts)
        return m_pimpl->fail("getSolverInputs: state, reference trajectory and contact phase list must be set");
    if (!m_pimpl->fillInputs(p, lbg, ubg, x0, nullptr)) return m_pimpl->fail("getSolverInputs: inconsistent contact phase list");
    return true;
}

bool CentroidalMPC::getTickRecord(double* tick) const
{
    if (m_pimpl->fsm == State::NotInitialized || !m_pimpl->hasState || !m_pimpl->hasReference || !m_pimpl->hasContacts)
        return m_pimpl->fail("getTickRecord: state, reference trajectory and contact phase list must be set");
    if (!m_pimpl->fillTick(tick)) return m_pimpl->fail("getTickRecord: inconsistent contact phase list");
    return true;
}

bool CentroidalMPC::advance() { return advanceBatch({this}); }

bool CentroidalMPC::advanceBatch(const std::vector<CentroidalMPC*>& controllers)
{
    if (controllers.empty()) return true;
    Impl& first = *controllers[0]->m_pimpl;
    const Layout L = first.L;
    const size_t B = controllers.size();
    for (auto* ctrl : controllers) {
        Impl& I = *ctrl->m_pimpl;
        I.fsm = I.fsm == State::NotInitialized ? State::NotInitialized : State::OutputInvalid;
        if (I.fsm == State::NotInitialized) return I.fail("advance: call initialize() first");
        if (!I.hasState || !I.hasReference || !I.hasContacts)
            return I.fail("advance: state, reference trajectory and contact phase list must be set");
        if (I.L.N != L.N || std::memcmp(&I.cfg, &first.cfg, sizeof(cmpc_config)) != 0)
            return I.fail("advanceBatch: the controllers must share one configuration");
    }
    if (!first.ensureHandle()) return false;
    const auto tic = std::chrono::steady_clock::now();
    // The host ships one compact tick record per controller (state, references, contact windows); the formal input
    // (p, lbg, ubg, x0) is expanded on the device (cmpc_populate) and the solution stays resident there for the warm start
    // of the next tick: per tick and controller 2.3 KB go up and x (plus lam_g with the warm start on) comes back.
    const int stride = cmpc_tick_stride(L.N);
    std::vector<double> T(B * stride), X(B * L.n());
    bool warm = true, wantLam = false;
    for (size_t b = 0; b < B; ++b) {
        Impl& I = *controllers[b]->m_pimpl;
        if (!I.fillTick(T.data() + b * stride)) return I.fail("advance: inconsistent contact phase list");
        warm = warm && I.warmStartEnabled && I.hasPrevious;
        wantLam = wantLam || I.warmStartEnabled;
    }
    std::vector<const CentroidalMPC*> batch(controllers.begin(), controllers.end());
    int warmMode = 0;
    std::vector<double> LAM;
    if (wantLam) LAM.assign(B * L.m(), 0.0);
    if (warm) {
        if (first.residentBatch == batch) warmMode = 1;   // the previous solutions of exactly this batch are on the device
        else {
            warmMode = 2;   // upload them: the device shifts them by one knot (cmpc_shift_warmstart)
            for (size_t b = 0; b < B; ++b) {
                Impl& I = *controllers[b]->m_pimpl;
                std::copy(I.xPrev.begin(), I.xPrev.end(), X.begin() + b * L.n());
                std::copy(I.lamPrev.begin(), I.lamPrev.end(), LAM.begin() + b * L.m());
            }
        }
    }
    std::vector<double> obj(B);
    std::vector<int> status(B), iters(B);
    const int rc = cmpc_solve_ticks_host(first.handle, (int)B, T.data(), warmMode, X.data(), wantLam ? LAM.data() : nullptr,
                                         obj.data(), status.data(), iters.data());
    if (rc != CMPC_OK) { first.residentBatch.clear(); return first.fail(std::string("cmpc_solve_ticks_host failed: ") + cmpc_error_string(rc)); }
    const auto elapsed = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tic);
    bool ok = true;
    for (size_t b = 0; b < B; ++b) {
        Impl& I = *controllers[b]->m_pimpl;
        I.x.assign(X.begin() + b * L.n(), X.begin() + (b + 1) * L.n());
        if (wantLam) I.lam.assign(LAM.begin() + b * L.m(), LAM.begin() + (b + 1) * L.m());
        I.stats.status = status[b]; I.stats.iterations = iters[b]; I.stats.objective = obj[b];
        // IPOPT failure => CasADi throws => BLF returns false; "Solved To Acceptable Level" counts as success for CasADi
        if (status[b] != CMPC_STATUS_CONVERGED && status[b] != CMPC_STATUS_ACCEPTABLE) {
            I.fail("advance: the solver did not converge (status " + std::to_string(status[b]) + ")");
            ok = false;
            continue;
        }
        I.unpack(elapsed);
        I.xPrev = I.x; I.lamPrev = I.lam; I.hasPrevious = true;
        I.currentTime += I.dT;  // the controller keeps its own clock: one sampling time per advance()
        I.fsm = State::OutputValid;
    }
    // a failed instance keeps its older previous solution on the host: what is resident no longer matches
    if (ok) first.residentBatch = batch; else first.residentBatch.clear();
    return ok;
}

const CentroidalMPCOutput& CentroidalMPC::getOutput() const { return m_pimpl->output; }
bool CentroidalMPC::isOutputValid() const { return m_pimpl->fsm == State::OutputValid; }
const CentroidalMPCSolverStats& CentroidalMPC::getSolverStats() const { return m_pimpl->stats; }
bool CentroidalMPC::getConfig(cmpc_config& cfg) const
{
    if (m_pimpl->fsm == State::NotInitialized) return false;
    cfg = m_pimpl->cfg;
    return true;
}
int CentroidalMPC::horizon() const { return m_pimpl->L.N; }
std::chrono::nanoseconds CentroidalMPC::samplingTime() const { return m_pimpl->dT; }
std::chrono::nanoseconds CentroidalMPC::currentTime() const { return m_pimpl->currentTime; }
const std::string& CentroidalMPC::lastError() const { return m_pimpl->error; }

}  // namespace ReducedModelControllers
}  // namespace BipedalLocomotion
