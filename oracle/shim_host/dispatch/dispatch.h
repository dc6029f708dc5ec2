/* oracle/shim_host/dispatch/dispatch.h -- TEST INFRASTRUCTURE.
 * Stand-in for the slice of libdispatch that RT_Metal/Metal/BVH.hh:30-270 uses. Work submitted with dispatch_async
 * runs immediately on the calling thread, so the reference's builder executes in its sequential order (left subtree,
 * then right subtree -- the variant the reference itself keeps as a commented call, BVH.hh:261) and its output is
 * deterministic. Blocks (`^{ ... }`) are turned into lambdas by the recipe in oracle/Makefile. */
#pragma once
#include <cstdint>

#define __block
#define DISPATCH_TIME_FOREVER (~0ull)
#define DISPATCH_QUEUE_CONCURRENT 0
#define DISPATCH_QUEUE_SERIAL 0

struct dispatch_queue_s {};
struct dispatch_group_s { int pending = 0; };
struct dispatch_semaphore_s { long value = 0; };
typedef dispatch_queue_s* dispatch_queue_t;
typedef dispatch_group_s* dispatch_group_t;
typedef dispatch_semaphore_s* dispatch_semaphore_t;

inline dispatch_queue_t dispatch_queue_create(const char*, int) { return new dispatch_queue_s(); }
inline dispatch_group_t dispatch_group_create() { return new dispatch_group_s(); }
inline dispatch_semaphore_t dispatch_semaphore_create(long v) { auto* s = new dispatch_semaphore_s(); s->value = v; return s; }
inline void dispatch_group_enter(dispatch_group_t g) { g->pending++; }
inline void dispatch_group_leave(dispatch_group_t g) { g->pending--; }
inline long dispatch_group_wait(dispatch_group_t, unsigned long long) { return 0; }
inline long dispatch_semaphore_wait(dispatch_semaphore_t, unsigned long long) { return 0; }
inline long dispatch_semaphore_signal(dispatch_semaphore_t) { return 0; }
template <typename F> inline void dispatch_async(dispatch_queue_t, F&& f) { f(); }
template <typename F> inline void dispatch_sync(dispatch_queue_t, F&& f) { f(); }
template <typename F> inline void dispatch_apply(unsigned long n, dispatch_queue_t, F&& f) { for (unsigned long i = 0; i < n; ++i) f(i); }
