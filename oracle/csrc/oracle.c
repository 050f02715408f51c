/* TEST INFRASTRUCTURE -- C helpers of the CPU oracle (oracle/port.py).  Never linked into the product.
 *
 * orc_cigar_scan   : analyze_cigar_indel restated (reference src/svim_asm/SVIM_intra.py:8-30) over
 *                    BAM-packed ops; used as the "C port" CPU baseline in bench.py.
 * orc_edit_distance: unit-cost global (Needleman-Wunsch) edit distance, the quantity the reference
 *                    obtains from edlib.align(a, b)["editDistance"] (SVIM_COMBINE.py:50,64,76,88,100;
 *                    edlib defaults mode="NW", task="distance" [ext]).  Plain two-row DP: it is the
 *                    independent check of the GPU's bit-parallel kernel, so it is deliberately simple.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* out: rows of 4 int64 (pos_ref, pos_read, length, is_del).  Returns the number of indels found
 * (may exceed cap; only the first cap rows are written). */
int64_t orc_cigar_scan(const uint32_t* ops, int64_t n_ops, int32_t min_length, int64_t* out, int64_t cap) {
    int64_t pos_ref = 0, pos_read = 0, found = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        const uint32_t op = ops[i] & 15u;
        const int64_t len = ops[i] >> 4;
        switch (op) {
            case 0: case 7: case 8:               /* :14-16, :27-29 */
                pos_ref += len;
                pos_read += len;
                break;
            case 1:                               /* :17-20 */
                if (len >= min_length) {
                    if (found < cap) { out[4 * found] = pos_ref; out[4 * found + 1] = pos_read; out[4 * found + 2] = len; out[4 * found + 3] = 0; }
                    ++found;
                }
                pos_read += len;
                break;
            case 2:                               /* :21-24 */
                if (len >= min_length) {
                    if (found < cap) { out[4 * found] = pos_ref; out[4 * found + 1] = pos_read; out[4 * found + 2] = len; out[4 * found + 3] = 1; }
                    ++found;
                }
                pos_ref += len;
                break;
            case 4:                               /* :25-26 */
                pos_read += len;
                break;
            default:                              /* N, H, P, B, pad: ignored */
                break;
        }
    }
    return found;
}

int64_t orc_edit_distance(const char* a, int64_t n, const char* b, int64_t m) {
    if (n < m) { const char* t = a; a = b; b = t; int64_t k = n; n = m; m = k; }
    if (m == 0) return n;
    int32_t* row = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    if (!row) return -1;
    for (int64_t j = 0; j <= m; ++j) row[j] = (int32_t)j;
    for (int64_t i = 1; i <= n; ++i) {
        int32_t diag = row[0];
        row[0] = (int32_t)i;
        const char ca = a[i - 1];
        for (int64_t j = 1; j <= m; ++j) {
            const int32_t up = row[j];
            int32_t best = diag + (ca != b[j - 1]);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            diag = up;
            row[j] = best;
        }
    }
    const int64_t d = row[m];
    free(row);
    return d;
}
