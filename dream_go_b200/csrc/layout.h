// Device memory layout of board activations ("board rows").
//
// Every tensor that a 3x3 convolution reads is stored as rows of C fp16 channels, one row per
// board point, in a halo-shared layout: point (x, y) of position n lives at row
//     m = n*400 + y*20 + x            (x, y in 0..18)
// Column x == 19 of every board line and the whole line y == 19 are halo rows that are always zero.
// The right halo of line y is the left halo of line y+1 and the bottom halo line of position n is
// the top halo line of position n+1, so a 3x3 tap (dy, dx) of ANY row m is simply row
// m + 20*dy + dx -- the convolution becomes nine row-shifted GEMMs over one contiguous buffer
// (the same trick as the reference board's 20-wide padded vertex array, src/libdg_go/point.rs:23-32).
// Padding overhead: 400/361 = 1.108 (a 21x21 zero-padded image would be 1.222).
//
// A buffer starts with DG_GUARD_ROWS zero rows (so row m is at buffer row DG_GUARD_ROWS + m and the
// -21 tap of row 0 stays in bounds) and ends with >= DG_TAIL_ROWS zero rows.
#pragma once

#define DG_LINE_STRIDE 20
#define DG_POS_ROWS 400
#define DG_HALO_ROWS 21          /* |20*dy + dx| <= 21 */
#define DG_GUARD_ROWS 32
#define DG_TAIL_ROWS 192
#define DG_TILE_M 128            /* rows per MMA tile (UMMA M) */
#define DG_WINDOW_ROWS (DG_TILE_M + 2 * DG_HALO_ROWS)   /* 170 rows staged per tile */

static inline int dg_num_tiles(int batch) { return (batch * DG_POS_ROWS + DG_TILE_M - 1) / DG_TILE_M; }
/* tiles are processed in pairs (one per CTA of a cluster), so buffers cover an even number of tiles */
static inline long dg_alloc_rows(int batch) { return (long)DG_GUARD_ROWS + (long)((dg_num_tiles(batch) + 1) & ~1) * DG_TILE_M + DG_TAIL_ROWS; }
