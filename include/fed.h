// Drop-in header: FED (Fast Explicit Diffusion) time-step schedule, same four entry points as the
// reference's fed.h:34-60.  Implemented over akz_fed_tau (akaze_b200.h).
#pragma once
#include <vector>

// the product library is built with -fvisibility=hidden: the drop-in surface is exported explicitly
#pragma GCC visibility push(default)

// tau receives the n step sizes of one cycle; returns n (0 on failure)
int fed_tau_by_process_time(const float T, const int M, const float tau_max, const bool reordering, std::vector<float>& tau);
int fed_tau_by_cycle_time(const float t, const float tau_max, const bool reordering, std::vector<float>& tau);
int fed_tau_internal(const int n, const float scale, const float tau_max, const bool reordering, std::vector<float>& tau);
bool fed_is_prime_internal(const int number);

#pragma GCC visibility pop
