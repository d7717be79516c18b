// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
// moveit_core collision_detection::AllowedCollisionMatrix, restated from its published behaviour for the four
// methods the reference calls (self_collision_model.cpp:285-307, 366-379, 1136, 1254): a symmetric map of
// name pairs -> NEVER / ALWAYS; setEntry(name, allowed) sets name against every KNOWN entry name;
// getEntry is false when the pair is absent.  (Same restatement as oracle/collision_model.h's ACM.)
#include <map>
#include <string>
#include <vector>
namespace collision_detection {
namespace AllowedCollision { enum Type { NEVER, ALWAYS, CONDITIONAL }; }
class AllowedCollisionMatrix
{
public:
    AllowedCollisionMatrix() { }
    template <typename Msg> explicit AllowedCollisionMatrix(const Msg& msg)   // moveit_msgs::AllowedCollisionMatrix
    {
        for (size_t i = 0; i < msg.entry_names.size(); ++i)
            for (size_t j = i; j < msg.entry_values.size() && j < msg.entry_values[i].enabled.size(); ++j)
                setEntry(msg.entry_names[i], msg.entry_names[j], msg.entry_values[i].enabled[j] != 0);
    }
    bool getEntry(const std::string& a, const std::string& b, AllowedCollision::Type& t) const
    {
        auto i = entries_.find(a);
        if (i == entries_.end()) return false;
        auto j = i->second.find(b);
        if (j == i->second.end()) return false;
        t = j->second;
        return true;
    }
    bool hasEntry(const std::string& a) const { return entries_.find(a) != entries_.end(); }
    bool hasEntry(const std::string& a, const std::string& b) const
    {
        auto i = entries_.find(a);
        return i != entries_.end() && i->second.find(b) != i->second.end();
    }
    void setEntry(const std::string& a, const std::string& b, bool allowed)
    {
        const AllowedCollision::Type v = allowed ? AllowedCollision::ALWAYS : AllowedCollision::NEVER;
        entries_[a][b] = entries_[b][a] = v;
    }
    void setEntry(const std::string& name, bool allowed)
    {
        std::string last = name;
        for (auto& e : entries_) {
            if (e.first != name) { last = e.first; setEntry(name, e.first, allowed); }
        }
    }
    void getAllEntryNames(std::vector<std::string>& names) const
    {
        names.clear();
        for (auto& e : entries_) names.push_back(e.first);
    }
    void clear() { entries_.clear(); }
private:
    std::map<std::string, std::map<std::string, AllowedCollision::Type>> entries_;
};
} // namespace collision_detection
