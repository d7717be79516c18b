// ORACLE (test infrastructure only).  See ../planners/planner.h (declarations of SBPL's DiscreteSpaceInformation).
#ifndef ORACLE_REF_STUBS_SBPL_ENVIRONMENT_H
#define ORACLE_REF_STUBS_SBPL_ENVIRONMENT_H
#include <sbpl/planners/planner.h>
#endif
