// dispatch.h -- the (n_elec, n_atoms) pairs with compiled instantiations.
// BASELINE.json configs: C all-electron (6,1), C ccECP (4,1), N2 ccECP (10,2), C6H6 ccECP (30,12);
// the reference's own examples: C2 ccECP (8,2), C2 all-electron (12,2); plus small test systems.
#pragma once
#ifdef AIQMC_SYSTEM_LIST   /* debugging aid: build.py AIQMC_SYSTEMS=... */
#define AIQMC_FOR_EACH_SYSTEM(X) AIQMC_SYSTEM_LIST
#else
#define AIQMC_FOR_EACH_SYSTEM(X) \
  X(2, 1) X(2, 2) X(3, 1) X(4, 1) X(4, 2) X(5, 2) X(6, 1) X(8, 2) X(10, 2) X(12, 2) X(30, 12)
#endif
