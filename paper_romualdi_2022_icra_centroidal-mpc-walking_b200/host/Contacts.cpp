// Contacts.cpp -- see BipedalLocomotion/Contacts.h
#include "BipedalLocomotion/Contacts.h"

#include <algorithm>

namespace BipedalLocomotion {
namespace Contacts {

bool ContactList::addContact(const PlannedContact& c)
{
    if (c.activationTime >= c.deactivationTime) return false;
    return m_contacts.insert(c).second;  // the comparator makes overlapping intervals "equivalent": insertion refused
}

bool ContactList::addContact(const manif::SE3d& pose, const std::chrono::nanoseconds& activationTime,
                             const std::chrono::nanoseconds& deactivationTime)
{
    PlannedContact c;
    c.pose = pose; c.activationTime = activationTime; c.deactivationTime = deactivationTime;
    c.name = m_name; c.index = m_index;
    return addContact(c);
}

bool ContactList::editContact(const_iterator it, const PlannedContact& c)
{
    if (it == m_contacts.end() || c.activationTime >= c.deactivationTime) return false;
    if (it != m_contacts.begin()) { auto p = it; --p; if (c.activationTime < p->deactivationTime) return false; }
    auto n = it; ++n;
    if (n != m_contacts.end() && c.deactivationTime > n->activationTime) return false;
    m_contacts.erase(it);
    return m_contacts.insert(c).second;
}

ContactList::const_iterator ContactList::getActiveContact(const std::chrono::nanoseconds& t) const
{
    for (auto it = m_contacts.begin(); it != m_contacts.end(); ++it)
        if (it->isContactActive(t)) return it;
    return m_contacts.end();
}

ContactList::const_iterator ContactList::getPresentContact(const std::chrono::nanoseconds& t) const
{
    auto best = m_contacts.end();
    for (auto it = m_contacts.begin(); it != m_contacts.end() && it->activationTime <= t; ++it) best = it;
    return best;
}

ContactList::const_iterator ContactList::getNextContact(const std::chrono::nanoseconds& t) const
{
    for (auto it = m_contacts.begin(); it != m_contacts.end(); ++it)
        if (it->activationTime > t) return it;
    return m_contacts.end();
}

bool ContactList::forceSampleTime(const std::chrono::nanoseconds& dT)
{
    if (dT <= std::chrono::nanoseconds::zero()) return false;
    std::set<PlannedContact, Compare> out;
    for (PlannedContact c : m_contacts) {
        c.activationTime -= c.activationTime % dT;
        if (c.deactivationTime != std::chrono::nanoseconds::max()) c.deactivationTime -= c.deactivationTime % dT;
        if (c.activationTime >= c.deactivationTime) return false;
        if (!out.insert(c).second) return false;
    }
    m_contacts.swap(out);
    return true;
}

bool ContactPhaseList::setLists(const ContactListMap& lists)
{
    m_lists = lists;
    buildPhases();
    return true;
}

void ContactPhaseList::buildPhases()
{
    m_phases.clear();
    std::vector<std::chrono::nanoseconds> ev;
    for (const auto& kv : m_lists)
        for (const auto& c : kv.second) { ev.push_back(c.activationTime); ev.push_back(c.deactivationTime); }
    std::sort(ev.begin(), ev.end());
    ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
        ContactPhase ph;
        ph.beginTime = ev[i]; ph.endTime = ev[i + 1];
        for (const auto& kv : m_lists) {
            auto it = kv.second.getActiveContact(ev[i]);
            if (it != kv.second.cend()) ph.activeContacts[kv.first] = it;
        }
        m_phases.push_back(ph);
    }
}

ContactPhaseList::const_iterator ContactPhaseList::getPresentPhase(const std::chrono::nanoseconds& t) const
{
    for (auto it = m_phases.begin(); it != m_phases.end(); ++it)
        if (t >= it->beginTime && t < it->endTime) return it;
    return m_phases.end();
}

bool ContactPhaseList::forceSampleTime(const std::chrono::nanoseconds& dT)
{
    for (auto& kv : m_lists)
        if (!kv.second.forceSampleTime(dT)) return false;
    buildPhases();
    return true;
}

}  // namespace Contacts
}  // namespace BipedalLocomotion
