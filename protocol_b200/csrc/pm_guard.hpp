// pm_guard.hpp — nothing unwinds across the C ABI (include/prime_match.h: "never unwind, never abort").
// Every extern "C" entry point that returns a pm_status is a function-try-block whose handler calls this
// Lippincott function; the Rust host maps the status to Err and retries next tick (reference convention:
// node_groups/scheduler_impl.rs:24-30, mod.rs:188-194).
#pragma once
#include <new>

#include "../../include/prime_match.h"

static inline int pm_guard_rc() noexcept {
  try {
    throw;
  } catch (const std::bad_alloc&) {
    return PM_E_NOMEM;
  } catch (...) {
    return PM_E_STATE;
  }
}
