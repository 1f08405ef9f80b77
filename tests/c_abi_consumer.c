/* A plain C99 consumer of include/sp_b200.h: what a non-Python, non-Julia host links against.  Built and run by
 * tests/test_abi_cpu.py::test_header_is_plain_c_and_a_c_program_links (no GPU needed: it exercises argument validation and
 * the "no CPU fallback" rule).  Exit code 0 = every expectation held. */
#include <stdio.h>
#include <string.h>

#include "sp_b200.h"

int main(void) {
    const double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {1.0, 1.0, 1.0};
    sp_system* sys = NULL;
    int32_t n = -1;
    if (sp_version() <= 0) return 10;
    if (sp_create(&sys, lo, hi, 0.0, 0) != SP_ERR_INVALID) return 11;          /* structs.jl:59: h must be positive */
    if (!strstr(sp_last_error(NULL), "h must be a positive float")) return 12;
    if (sp_create(NULL, lo, hi, 0.1, 0) != SP_ERR_INVALID) return 13;
    if (sp_device_count(&n) != SP_OK || n == 0) {
        /* no GPU: creation fails loudly */
        if (sp_create(&sys, lo, hi, 0.1, 0) != SP_ERR_NO_DEVICE) return 14;
        if (sys != NULL) return 15;
        printf("c-abi consumer ok (no device)\n");
        return 0;
    }
    /* with a GPU: one tiny system through the whole life cycle */
    if (sp_create(&sys, lo, hi, 0.1, 0) != SP_OK) return 16;
    {
        const double x[6] = {0.25, 0.75, 0.5, 0.5, 0.5, 0.5}; /* SoA: x0 x1 | y0 y1 | z0 z1 */
        int64_t count = -1;
        if (sp_resize(sys, 2) != SP_OK) return 17;
        if (sp_upload(sys, 0, x, 2, SP_LAYOUT_SOA) != SP_OK) return 18; /* field 0 is the position */
        if (sp_create_cell_list(sys) != SP_OK) return 19;
        if (sp_num_particles(sys, &count) != SP_OK || count != 2) return 20;
    }
    if (sp_destroy(sys) != SP_OK) return 21;
    printf("c-abi consumer ok (device)\n");
    return 0;
}
