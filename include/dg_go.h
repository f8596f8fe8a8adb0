/*
 * dg_go.h -- C ABI of the host half of the self-play hot path: Go rules, legal-move generation, V1 feature
 * extraction into the engine's compact position format, symmetry tables and prior construction.
 *
 * Replaces, for this path, the Rust crate API of `dg_go` / `dg_mcts::pool::policy_helper`
 * (SURVEY.md section 8 rows a1-a3, a14).  The functions are exported by the same `libdg_engine.so` as
 * include/dg_engine.h.  Points are the packed indices the network uses, 19*y + x with (x, y) as in
 * `Point::new(x, y)` (src/libdg_go/point.rs:26-34,137-143); 361 = pass.  Colours: 1 = black, 2 = white
 * (src/libdg_go/color.rs:18-21).  Symmetries are numbered in the order of `symmetry::ALL`
 * (src/libdg_go/utils/symmetry.rs:121-130): Identity, FlipLR, FlipUD, Transpose, TransposeAnti, Rot90,
 * Rot180, Rot270.
 */
#ifndef DG_GO_H
#define DG_GO_H

#include <stdint.h>
#include "dg_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DG_BLACK 1
#define DG_WHITE 2
#define DG_PASS  361

typedef struct dg_board dg_board;

/* Out-of-range arguments (point outside 0..360, colour outside 1..2, transform outside 0..7, unknown search kind) are
 * rejected at this boundary: queries answer 0 / -1, dg_board_place does nothing -- in particular for 361, the pass index
 * the search hands out as a move (the reference does not call `place` for a pass, self_play.rs:442-451) --, functions
 * that fill an array fill it with zeros (no plane, no legal move, no candidate; dg_board_prior: all -inf), dg_go_replay
 * stops at the ply as at an illegal move.  Pointers must be valid (board handles, output arrays of the stated size). */

/* The zobrist constants of the process, in the reference's layout `zobrist::TABLE: [[u64; 420]; 3]` (src/libdg_go/
 * zobrist.rs:18; index 20 * (y + 1) + (x + 1), board_fast.rs).  The shim passes `&dg_go::zobrist::TABLE` once at start-up
 * and every hash -- `Board::zobrist_hash`, the super-ko window, the transposition-table key, the hashes the device feature
 * kernel compares -- is then the reference's (dg_tests/tests/real_games.rs:49,74,117).  NULL restores the built-in
 * table.  Call before creating boards and engines (an engine uploads the table when it is created); not thread-safe. */
void      dg_go_set_zobrist(const uint64_t* table /* [3][420] or NULL */);

/* ---- `Board` (src/libdg_go/board.rs) ------------------------------------------------------------------ */
dg_board* dg_board_new(float komi);                                   /* Board::new          board.rs:51-62   */
dg_board* dg_board_clone(const dg_board* board);                      /* #[derive(Clone)]    board.rs:26      */
void      dg_board_copy(dg_board* dst, const dg_board* src);
void      dg_board_free(dg_board* board);
void      dg_board_set_komi(dg_board* board, float komi);             /* board.rs:78-81 */
float     dg_board_komi(const dg_board* board);
int32_t   dg_board_count(const dg_board* board);                      /* board.rs:84-87 */
uint64_t  dg_board_zobrist_hash(const dg_board* board);               /* board.rs:90-93; constants: dg_go_set_zobrist */
int32_t   dg_board_to_move(const dg_board* board);                    /* board.rs:102-107 */
int32_t   dg_board_at(const dg_board* board, int32_t point);          /* board.rs:117-121; 0 = empty */
/* Board::is_valid (board.rs:151-153): empty, not suicide, not a repetition of the last 16 positions. */
int32_t   dg_board_is_valid(const dg_board* board, int32_t color, int32_t point);
/* Board::place (board.rs:164-188): plays without checking legality; captures, history, hash. */
void      dg_board_place(dg_board* board, int32_t color, int32_t point);
/* BoardFast::get_n_liberty / get_n_liberty_if (board_fast.rs:170-174, 484-539); the latter is -1 for an illegal move. */
int32_t   dg_board_get_n_liberty(const dg_board* board, int32_t point);
int32_t   dg_board_get_n_liberty_if(const dg_board* board, int32_t color, int32_t point);
/* Ladder::is_ladder_capture / is_ladder_escape (utils/ladder.rs:131-178); `point` must be a legal move. */
int32_t   dg_board_is_ladder_capture(const dg_board* board, int32_t color, int32_t point);
int32_t   dg_board_is_ladder_escape(const dg_board* board, int32_t color, int32_t point);
/* symmetry::is_symmetric (utils/symmetry.rs:139-146) */
int32_t   dg_board_is_symmetric(const dg_board* board, int32_t transform);
/* 361 x Board::is_valid -- the legal-move generation of pool/policy_helper.rs:39-43 (StandardSearch). */
void      dg_board_legal_moves(const dg_board* board, int32_t color, uint8_t* out /* [361] */);

/* ---- symmetry (src/libdg_go/utils/symmetry.rs) ------------------------------------------------------------ */
int32_t   dg_symmetry_apply(int32_t transform, int32_t point);        /* Transform::apply  :91-102 */
int32_t   dg_symmetry_inverse(int32_t transform);                     /* Transform::inverse :78-89 */

/* ---- `features::V1::get_features` (src/libdg_go/utils/features.rs:154-250) ----------------------------------- */
/* Compact form (what the engine's queue / dg_engine_forward_packed takes).  `legal` (optional, [361]) receives
 * Board::is_valid(to_move, .) in identity orientation -- computed from the same pass at no extra cost. */
void      dg_board_features_packed(const dg_board* board, int32_t to_move, int32_t symmetry,
                                   dg_packed_position* out, uint8_t* legal);
/* The raw form for dg_engine_forward_raw: stones, visited bits, hashes, last moves and the two ladder planes; the
 * device derives the other 30 planes and the legal moves from it (csrc/features.cu).  `symmetry` may carry the search
 * options of the position in bits 4.. (DG_SCORING_SEARCH << 4) for dg_engine_forward_raw_prior, and DG_RAW_DEVICE_LADDERS
 * (0x08): the ladders are not read here, the device reads them. */
void      dg_board_raw_position(const dg_board* board, int32_t to_move, int32_t symmetry, dg_raw_position* out);
/* Bit-identical to `get_features::<HWC, f16>`: 11,552 fp16, index 32*(19y+x)+c. */
void      dg_board_features_f16(const dg_board* board, int32_t to_move, int32_t symmetry, uint16_t* out);
/* The same for `count` boards on up to `threads` host threads (<= 0: all cores); this is BASELINE.json configs[0]
 * (feature-plane extract + legal-move generation for a batch of boards). */
void      dg_go_extract_batch(const dg_board* const* boards, const uint8_t* to_move, const uint8_t* symmetry,
                              int32_t count, dg_packed_position* out, uint8_t* legal /* [count][361] or NULL */,
                              int32_t threads);

/* Replays `n` moves of one game (colour, point; point 361 = pass, not played -- self_play.rs:442-451) and writes,
 * for the position BEFORE each move, the compact features (identity symmetry, to_move = the move's colour), the
 * legal mask and / or the hash AFTER the move (any may be NULL).  Returns n, or -(ply+1) at the first illegal move
 * (dg_tests/tests/common/mod.rs:60). */
int32_t   dg_go_replay(float komi, const uint8_t* colors, const uint16_t* moves, int32_t n,
                       dg_packed_position* features, uint8_t* legal, uint64_t* hashes);

/* ---- prior construction (src/libdg_mcts/pool/policy_helper.rs, predictor.rs:30-44) --------------------------- */
/* create_initial_policy (for the given `search` options) + add_valid_candidates + normalize_policy(sum_to), as the Insert event of
 * pool/worker_thread.rs:88-93 runs them: `policy` is the network output (362 fp16, in the orientation `symmetry`
 * the features were extracted with); `prior` receives 368 floats (-inf = not a candidate, 362..367 padding = -inf).
 * `legal` may be NULL (recomputed) or the mask dg_board_features_packed returned. */
void      dg_board_prior(const dg_board* board, int32_t to_move, int32_t search, const uint8_t* legal, const uint16_t* policy,
                         int32_t symmetry, float sum_to, float* prior /* [368] */);

/* ---- what self-play needs on top (utils/benson.rs, utils/score.rs, libdg_mcts/options.rs) --------------------- */
#define DG_STANDARD_SEARCH 0   /* StandardSearch: pass or Board::is_valid                      options.rs:53-57   */
#define DG_SCORING_SEARCH  1   /* ScoringSearch: no pass, no Benson eye of either colour, no simple own eye  :109-138 */
/* Score::is_scorable (utils/score.rs:97-110) */
int32_t   dg_board_is_scorable(const dg_board* board);
/* Benson status per point for `color`: 0 none, 1 unconditionally alive stone, 2 vital region (benson.rs:152-165) */
void      dg_board_benson(const dg_board* board, int32_t color, uint8_t* out /* [361] */);
/* Whose territory the game record counts each point as (utils/score.rs:148-195, game_result.rs:45-93): 1 black,
 * 2 white, 0 neither -- `RE[]` and `TB[]/TW[]` of a finished game. */
void      dg_board_territory(const dg_board* board, uint8_t* out /* [361] */);
/* PolicyChecker::is_policy_candidate for moves 0..361 under `search`; `legal` optional as above. */
void      dg_board_policy_candidates(const dg_board* board, int32_t to_move, int32_t search, const uint8_t* legal,
                                     uint8_t* out /* [362] */);

#ifdef __cplusplus
}
#endif
#endif /* DG_GO_H */
