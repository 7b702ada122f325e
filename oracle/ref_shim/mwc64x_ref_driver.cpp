// mwc64x_ref_driver.cpp -- runs the reference's own MWC64X kernels on the host, one
// "work-item" at a time.  TEST INFRASTRUCTURE; compiled only into oracle/_ref/.
#include "opencl_shim.h"
// the sed-translated reference sources (see oracle/Makefile)
#include "randstategen.cl.h"
#include "randomnumbergenerator.cl.h"

extern "C" {
__attribute__((visibility("default"))) void ref_generate_random_state(uint32_t* seeds, int size) {
    for (int i = 0; i < size; ++i) {
        g_global_id = (size_t)i;
        MWC64X_GenerateRandomState(seeds, size);
    }
}
__attribute__((visibility("default"))) void ref_generate_per_stream_random_state(uint32_t* seeds, uint64_t gap, int size) {
    for (int i = 0; i < size; ++i) {
        g_global_id = (size_t)i;
        MWC64X_GeneratePerStreamRandomState(seeds, gap, size);
    }
}
__attribute__((visibility("default"))) void ref_random_number_generator(uint32_t* seeds, int size, float* out) {
    for (int i = 0; i < size; ++i) {
        g_global_id = (size_t)i;
        randomNumberGeneratorKernel(seeds, size, out);
    }
}
__attribute__((visibility("default"))) void ref_step(uint32_t* x, uint32_t* c) {
    random_state s; s.x = *x; s.c = *c;
    MWC64X_Step(&s);
    *x = s.x; *c = s.c;
}
}
