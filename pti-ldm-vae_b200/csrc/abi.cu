#include "../../include/ptivae.h"
extern "C" int ptivae_abi_version(void) { return 100 * 1000 + 1; }  // sm_100a, ABI rev 1
