/* A plain C99 consumer of include/rrin_b200.h: what a maintainer's C / cgo / JNI stub would compile against.
 * Host-side entry points only (no GPU needed): version, the conv table, engine planning and its error paths.
 * Built and run by tests/test_host_logic.py::test_c99_consumer_links_and_runs. */
#include <stdio.h>
#include <string.h>
#include "rrin_b200.h"

#define CHECK(cond) do { if (!(cond)) { printf("FAILED %s (line %d): %s\n", #cond, __LINE__, rrin_last_error()); return 1; } } while (0)

int main(void) {
    char key[96];
    int cin, cout, level, src_mode, act, i, n, heads = 0, lasts = 0;
    rrin_engine* e = NULL;
    size_t ws;

    CHECK(rrin_version() > 0);
    n = rrin_num_convs();
    CHECK(n == 81);                                                    /* 23 + 3 x 19 + ... : unet.py:24-38 at depth 5 / 4 */
    for (i = 0; i < n; ++i) {
        CHECK(rrin_conv_info(i, key, (int)sizeof key, &cin, &cout, &level, &src_mode, &act) == RRIN_OK);
        CHECK(cin > 0 && cout > 0 && level >= 0 && level <= 4);
        if (src_mode == 4) ++heads;
        if (strlen(key) > 5 && strcmp(key + strlen(key) - 5, ".last") == 0) { ++lasts; CHECK(act == 0); }
    }
    CHECK(heads == 4 && lasts == 4);                                    /* one head and one `last` per U-Net, model.py:27-30 */
    CHECK(rrin_conv_info(n, key, (int)sizeof key, &cin, &cout, &level, &src_mode, &act) != RRIN_OK);
    CHECK(rrin_packed_weights_bytes() > 19194445u * 2u);               /* >= the 19.2 M parameters in 16 bits */

    CHECK(rrin_engine_create(1, 1, 72, 80, &e) == RRIN_ERR_BAD_SHAPE);  /* 72 % 16 != 0 */
    CHECK(strstr(rrin_last_error(), "16") != NULL);
    CHECK(rrin_engine_create(4, 4, 1088, 1920, &e) == RRIN_OK && e != NULL);
    ws = rrin_engine_workspace_bytes(e);
    CHECK(ws > 0 && ws % 256 == 0);
    CHECK(rrin_engine_num_launches(e) >= 81);
    printf("version %d convs %d launches %d workspace %zu\n", rrin_version(), n, rrin_engine_num_launches(e), ws);
    rrin_engine_destroy(e);
    return 0;
}
