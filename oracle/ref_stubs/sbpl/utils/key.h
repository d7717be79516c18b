// ORACLE (test infrastructure only).  The reference's arastar.cpp includes <sbpl/utils/key.h> but uses
// nothing from it.
#ifndef ORACLE_REF_STUBS_SBPL_KEY_H
#define ORACLE_REF_STUBS_SBPL_KEY_H
#endif
