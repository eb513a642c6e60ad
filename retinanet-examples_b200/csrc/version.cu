// version.cu -- library identification.
#include "common.cuh"
extern "C" const char *odtk_b200_version(void) { return "odtk_b200 0.1 sm_100a"; }
