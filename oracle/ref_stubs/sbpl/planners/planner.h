// ORACLE (test infrastructure only -- never linked into the product path).
//
// Minimal restatement of the part of SBPL's public planner interface that the REFERENCE's
// smpl/src/search/arastar.cpp is compiled against (sbpl/planners/planner.h, sbpl/discrete_space_information/
// environment.h of the Search-Based Planning Library, a third-party dependency that is not under
// /root/reference and not installed here).  Only declarations the reference's ARA* names are present; there
// is no SBPL code in this file.  Used by `make -C oracle ref` to build oracle/_ref/libref_arastar.so.
#ifndef ORACLE_REF_STUBS_SBPL_PLANNER_H
#define ORACLE_REF_STUBS_SBPL_PLANNER_H

#include <vector>

#define INFINITECOST 1000000000

#include <cstdio>

#define NUMOFINDICES_STATEID2IND 2

struct MDPConfig { int startstateid; int goalstateid; };
class CMDPSTATE;

// The virtuals RobotPlanningSpace overrides (smpl/graph/robot_planning_space.h:115-180, 211-214): SBPL's
// DiscreteSpaceInformation plus the ...ByGroup / ...WithExpansion entries of the author's SBPL fork.  Declarations
// only; StateID2IndexMapping is the per-state scratch array SBPL planners index by state id.
class DiscreteSpaceInformation
{
public:
    std::vector<int*> StateID2IndexMapping;

    virtual ~DiscreteSpaceInformation() { for (int* p : StateID2IndexMapping) delete[] p; }
    virtual bool InitializeEnv(const char*) { return false; }
    virtual bool InitializeMDPCfg(MDPConfig*) { return false; }
    virtual int GetFromToHeuristic(int, int) { return 0; }
    virtual int GetGoalHeuristic(int) { return 0; }
    virtual int GetStartHeuristic(int) { return 0; }
    virtual void GetSuccs(int SourceStateID, std::vector<int>* SuccIDV, std::vector<int>* CostV) = 0;
    virtual void GetPreds(int, std::vector<int>*, std::vector<int>*) { }
    virtual void GetLazySuccs(int, std::vector<int>*, std::vector<int>*, std::vector<bool>*) { }
    virtual int GetTrueCost(int, int) { return -1; }
    virtual void GetSuccsByGroup(int, std::vector<int>*, std::vector<int>*, std::vector<int>*, int) { }
    virtual void GetSuccsByGroupAndExpansion(int, std::vector<int>*, std::vector<int>*, int, int) { }
    virtual void GetSuccsWithExpansion(int, std::vector<int>*, std::vector<int>*, int) { }
    virtual void GetPredsByGroupAndExpansion(int, std::vector<int>*, std::vector<int>*, std::vector<int>*, int, int, int) { }
    virtual bool updateMultipleStartStates(std::vector<int>*, std::vector<double>*, int) { return false; }
    virtual void SetAllActionsandAllOutcomes(CMDPSTATE*) { }
    virtual int SizeofCreatedEnv() { return 0; }
    virtual void PrintState(int, bool, FILE* = nullptr) { }
};

class StateChangeQuery
{
public:
    virtual ~StateChangeQuery() { }
    virtual std::vector<int> const* getPredecessors() const = 0;
    virtual std::vector<int> const* getSuccessors() const = 0;
};

struct PlannerStats
{
    double eps;
    int cost;
    double time;
    int expands;
};

class ReplanParams
{
public:
    ReplanParams(double time)
    {
        max_time = time;
        initial_eps = 5.0;
        final_eps = 1.0;
        dec_eps = 0.2;
        return_first_solution = false;
        repair_time = -1;
    }
    double initial_eps, final_eps, dec_eps;
    bool return_first_solution;
    double max_time, repair_time;
};

class SBPLPlanner
{
public:
    virtual ~SBPLPlanner() { }
    virtual int replan(double allocated_time_sec, std::vector<int>* solution_stateIDs_V) = 0;
    virtual int replan(double allocated_time_sec, std::vector<int>* solution_stateIDs_V, int* solcost) = 0;
    virtual int replan(std::vector<int>* solution_stateIDs_V, ReplanParams params) { return 0; }
    virtual int replan(std::vector<int>* solution_stateIDs_V, ReplanParams params, int* solcost) { return 0; }
    virtual int set_goal(int goal_stateID) = 0;
    virtual int set_start(int start_stateID) = 0;
    virtual int force_planning_from_scratch() = 0;
    virtual int force_planning_from_scratch_and_free_memory() { return 0; }
    virtual int set_search_mode(bool bSearchUntilFirstSolution) = 0;
    virtual void costs_changed(StateChangeQuery const& stateChange) = 0;
    virtual double get_solution_eps() const { return -1; }
    virtual int get_n_expands() const { return -1; }
    virtual double get_initial_eps() { return -1; }
    virtual double get_initial_eps_planning_time() { return -1; }
    virtual double get_final_eps_planning_time() { return -1; }
    virtual int get_n_expands_init_solution() { return -1; }
    virtual double get_final_epsilon() { return -1; }
    virtual void get_search_stats(std::vector<PlannerStats>* s) { }
    virtual void set_initialsolution_eps(double initialsolution_eps) { }

protected:
    DiscreteSpaceInformation* environment_;
    bool bforwardsearch;
};

#endif
