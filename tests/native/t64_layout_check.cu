// Host-only check of the T64 corpus layout (persian-rag-system_b200/csrc/common.cuh::t64_offset):
// within every 64-row block the mapping (row, 16-byte chunk) -> byte offset must be a bijection onto
// the block, k-block-major, with the 128-byte-swizzle pattern tcgen05 expects (chunk ^ (row & 7)).
#include <cstdio>
#include <vector>
#include "../../persian-rag-system_b200/csrc/common.cuh"

int main() {
    using namespace prs;
    for (int pitch : {64, 128, 384, 512, 768}) {
        const int chunks = pitch / 8;                       // 16-byte chunks per row
        const size_t blk_bytes = (size_t)BLK_ROWS * pitch * 2;
        for (long long blk : {0ll, 1ll, 7ll, 123456ll}) {
            std::vector<char> seen(blk_bytes / 16, 0);
            for (int r = 0; r < BLK_ROWS; ++r) {
                for (int c = 0; c < chunks; ++c) {
                    const size_t off = t64_offset(blk * BLK_ROWS + r, c, pitch);
                    if (off < (size_t)blk * blk_bytes || off >= (size_t)(blk + 1) * blk_bytes || off % 16) { printf("out of block\n"); return 1; }
                    const size_t local = off - (size_t)blk * blk_bytes;
                    if (seen[local / 16]++) { printf("collision\n"); return 1; }
                    // k-block major: chunk c belongs to k-block c/8, which is an 8 KB piece [64 rows][128 B]
                    if (local / KBLOCK_BYTES != (size_t)(c / 8)) { printf("not k-block major\n"); return 1; }
                    const size_t in_piece = local % KBLOCK_BYTES;
                    if (in_piece / 128 != (size_t)r) { printf("row pieces are not 128 bytes apart\n"); return 1; }
                    if ((in_piece % 128) / 16 != (size_t)((c % 8) ^ (r % 8))) { printf("swizzle mismatch\n"); return 1; }
                }
            }
            for (char s : seen) if (s != 1) { printf("hole\n"); return 1; }
        }
    }
    printf("ok\n");
    return 0;
}
