/* oracle/shim/log/log.hpp -- TEST INFRASTRUCTURE ONLY.  No-op stand-in for mika314/log
 * (un-vendored dependency of the reference; spec.cpp includes it but never calls LOG). */
#pragma once
#define LOG(...) do { } while (0)
