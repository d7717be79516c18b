// ORACLE (test infrastructure only -- never linked into the product path).  Stand-in for an absent third-party header,
// just enough surface for the reference's sbpl_collision_checking sources to COMPILE where they lie (make -C oracle ref ->
// oracle/_ref/libref_collision.so).  No behaviour of the hot path lives here unless the header says so.
#ifndef STUB_CONSOLE_BRIDGE_H
#define STUB_CONSOLE_BRIDGE_H
#define logDebug(...) do { } while (0)
#define logInform(...) do { } while (0)
#define logWarn(...) do { } while (0)
#define logError(...) do { } while (0)
#define CONSOLE_BRIDGE_logDebug(...) do { } while (0)
#define CONSOLE_BRIDGE_logInform(...) do { } while (0)
#define CONSOLE_BRIDGE_logWarn(...) do { } while (0)
#define CONSOLE_BRIDGE_logError(...) do { } while (0)
#endif
