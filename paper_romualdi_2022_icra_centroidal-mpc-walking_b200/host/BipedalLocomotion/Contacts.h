// Contacts.h -- the BLF Contacts data model at the boundary of CentroidalMPC, restated from the way the reference uses it
// (src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:32-110: lists(), getNextContact, getActiveContact, cend, addContact,
//  setLists; :588 forceSampleTime; src/WholeBodyQPBlock.cpp:824-828, 1092, 1319-1330: DiscreteGeometryContact::corners
//  {position, force}, pose).  Times are std::chrono::nanoseconds as in the reference.  BLF itself is an un-vendored
// dependency: semantics marked [RECALL] follow its public documentation.
#pragma once

#include <chrono>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "BipedalLocomotion/Math.h"

namespace BipedalLocomotion {
namespace Contacts {

enum class ContactType { FULL, POINT };

struct ContactBase {
    manif::SE3d pose;
    std::string name{"Contact"};
    int index{-1};
    ContactType type{ContactType::FULL};
};

struct PlannedContact : public ContactBase {
    std::chrono::nanoseconds activationTime{std::chrono::nanoseconds::zero()};
    std::chrono::nanoseconds deactivationTime{std::chrono::nanoseconds::max()};
    // active in [activationTime, deactivationTime)
    bool isContactActive(const std::chrono::nanoseconds& t) const { return t >= activationTime && t < deactivationTime; }
};

struct Corner {
    Eigen::Vector3d position;  // in the contact frame
    Eigen::Vector3d force;     // in the inertial frame, per unit of robot mass on the MPC path
};

struct DiscreteGeometryContact : public ContactBase {
    std::vector<Corner> corners;
};

class ContactList {
    struct Compare {
        bool operator()(const PlannedContact& a, const PlannedContact& b) const { return a.deactivationTime <= b.activationTime; }
    };
public:
    using const_iterator = std::set<PlannedContact, Compare>::const_iterator;

    void setDefaultName(const std::string& n) { m_name = n; }
    const std::string& defaultName() const { return m_name; }
    void setDefaultIndex(int i) { m_index = i; }
    int defaultIndex() const { return m_index; }

    // false when the activation interval is empty or overlaps a contact already in the list
    bool addContact(const PlannedContact& c);
    bool addContact(const manif::SE3d& pose, const std::chrono::nanoseconds& activationTime,
                    const std::chrono::nanoseconds& deactivationTime);
    // replace the contact at `it` (the new interval must not overlap its neighbours)
    bool editContact(const_iterator it, const PlannedContact& c);
    const_iterator erase(const_iterator it) { return m_contacts.erase(it); }
    void clear() { m_contacts.clear(); }
    size_t size() const { return m_contacts.size(); }

    const_iterator begin() const { return m_contacts.begin(); }
    const_iterator end() const { return m_contacts.end(); }
    const_iterator cbegin() const { return m_contacts.cbegin(); }
    const_iterator cend() const { return m_contacts.cend(); }
    const_iterator firstContact() const { return m_contacts.begin(); }
    const_iterator lastContact() const { return m_contacts.empty() ? m_contacts.end() : --m_contacts.end(); }
    // the contact active at t, cend() when the foot is in the air
    const_iterator getActiveContact(const std::chrono::nanoseconds& t) const;
    // the last contact whose activation time is <= t, cend() when there is none
    const_iterator getPresentContact(const std::chrono::nanoseconds& t) const;
    // the first contact whose activation time is > t, cend() when there is none
    const_iterator getNextContact(const std::chrono::nanoseconds& t) const;
    // activation / deactivation times moved down to multiples of dT [RECALL BLF ContactList::forceSampleTime]
    bool forceSampleTime(const std::chrono::nanoseconds& dT);

private:
    std::set<PlannedContact, Compare> m_contacts;
    std::string m_name{"ContactList"};
    int m_index{-1};
};

using ContactListMap = std::map<std::string, ContactList>;

struct ContactPhase {
    std::chrono::nanoseconds beginTime{0}, endTime{0};
    std::map<std::string, ContactList::const_iterator> activeContacts;
};

class ContactPhaseList {
public:
    using const_iterator = std::vector<ContactPhase>::const_iterator;
    ContactPhaseList() = default;
    // the phases hold iterators into the lists: copies rebuild them
    ContactPhaseList(const ContactPhaseList& o) : m_lists(o.m_lists) { buildPhases(); }
    ContactPhaseList& operator=(const ContactPhaseList& o) { if (this != &o) { m_lists = o.m_lists; buildPhases(); } return *this; }
    bool setLists(const ContactListMap& lists);
    const ContactListMap& lists() const { return m_lists; }
    const_iterator begin() const { return m_phases.begin(); }
    const_iterator end() const { return m_phases.end(); }
    const_iterator cbegin() const { return m_phases.cbegin(); }
    const_iterator cend() const { return m_phases.cend(); }
    size_t size() const { return m_phases.size(); }
    const_iterator firstPhase() const { return m_phases.begin(); }
    const_iterator lastPhase() const { return m_phases.empty() ? m_phases.end() : --m_phases.end(); }
    const_iterator getPresentPhase(const std::chrono::nanoseconds& t) const;
    bool forceSampleTime(const std::chrono::nanoseconds& dT);
    void clear() { m_lists.clear(); m_phases.clear(); }

private:
    void buildPhases();
    ContactListMap m_lists;
    std::vector<ContactPhase> m_phases;
};

}  // namespace Contacts
}  // namespace BipedalLocomotion
