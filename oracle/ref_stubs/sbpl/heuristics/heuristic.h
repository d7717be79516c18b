// ORACLE (test infrastructure only).  Minimal restatement of SBPL's Heuristic interface
// (sbpl/heuristics/heuristic.h); see ../planners/planner.h.
#ifndef ORACLE_REF_STUBS_SBPL_HEURISTIC_H
#define ORACLE_REF_STUBS_SBPL_HEURISTIC_H

#include <sbpl/planners/planner.h>

class Heuristic
{
public:
    Heuristic(DiscreteSpaceInformation* environment) : m_environment(environment) { }
    virtual ~Heuristic() { }
    virtual int GetGoalHeuristic(int state_id) = 0;
    virtual int GetStartHeuristic(int state_id) = 0;
    virtual int GetFromToHeuristic(int from_id, int to_id) = 0;

protected:
    DiscreteSpaceInformation* m_environment;
};

#endif
