// TEST INFRASTRUCTURE ONLY (oracle/): pins the entropy source of the UNMODIFIED reference. The reference seeds
// a fresh std::mt19937 from std::random_device before every sweep of its chinese-whispers loops
// (src/cluster_graph.cpp:255-258,429-432), so its .gro output differs from run to run. Linking this object into
// the reference executable (or, with -Bsymbolic, into the shim library) replaces libstdc++'s out-of-line
// std::random_device::_M_getval() by a constant: every shuffle of n elements is then the same permutation, and
// the product's HS_PIN_SEED=<the same constant> mode must reproduce the reference's output byte for byte.
#include <random>

#ifndef HS_PIN_VALUE
#define HS_PIN_VALUE 20260117u
#endif

std::random_device::result_type std::random_device::_M_getval() { return HS_PIN_VALUE; }
