/* tetris_b200_features.h — feature ids of libtetris_b200.so <-> names of cogworks/tetris/features.py.
 *
 * id = position of the feature in the reference file (source order, features.py:18-347); the id is what
 * tb200_set_features() takes and the column index of tb200_eval_features()'s f64[N][45] output.
 * tests/test_host_logic.py checks it against tetris_ai_b200/tetris/features.py and the oracle's table.
 */
#ifndef TETRIS_B200_FEATURES_H
#define TETRIS_B200_FEATURES_H

enum tb200_feature_id {
    TB200_F_MEAN_HT          =  0,   /* features.py:25 */
    TB200_F_MAX_HT           =  1,   /* features.py:31 */
    TB200_F_MIN_HT           =  2,   /* features.py:37 */
    TB200_F_ALL_HT           =  3,   /* features.py:43 */
    TB200_F_MAX_HT_DIFF      =  4,   /* features.py:49 */
    TB200_F_MIN_HT_DIFF      =  5,   /* features.py:55 */
    TB200_F_CLEARED          =  6,   /* features.py:63 */
    TB200_F_CUML_CLEARED     =  7,   /* features.py:70 */
    TB200_F_TETRIS           =  8,   /* features.py:76 */
    TB200_F_COL_TRANS        =  9,   /* features.py:84 */
    TB200_F_ROW_TRANS        = 10,   /* features.py:90 */
    TB200_F_ALL_TRANS        = 11,   /* features.py:97 */
    TB200_F_WELLS            = 12,   /* features.py:117 */
    TB200_F_CUML_WELLS       = 13,   /* features.py:124 */
    TB200_F_DEEP_WELLS       = 14,   /* features.py:130 */
    TB200_F_MAX_WELL         = 15,   /* features.py:136 */
    TB200_F_PITS             = 16,   /* features.py:159 */
    TB200_F_PIT_ROWS         = 17,   /* features.py:165 */
    TB200_F_LUMPED_PITS      = 18,   /* features.py:172 */
    TB200_F_PIT_DEPTH        = 19,   /* features.py:183 */
    TB200_F_MEAN_PIT_DEPTH   = 20,   /* features.py:193 */
    TB200_F_CD_1             = 21,   /* features.py:207 */
    TB200_F_CD_2             = 22,   /* features.py:208 */
    TB200_F_CD_3             = 23,   /* features.py:209 */
    TB200_F_CD_4             = 24,   /* features.py:210 */
    TB200_F_CD_5             = 25,   /* features.py:211 */
    TB200_F_CD_6             = 26,   /* features.py:212 */
    TB200_F_CD_7             = 27,   /* features.py:213 */
    TB200_F_CD_8             = 28,   /* features.py:214 */
    TB200_F_CD_9             = 29,   /* features.py:215 */
    TB200_F_ALL_DIFFS        = 30,   /* features.py:220 */
    TB200_F_MAX_DIFFS        = 31,   /* features.py:226 */
    TB200_F_JAGGEDNESS       = 32,   /* features.py:232 */
    TB200_F_D_MAX_HT         = 33,   /* features.py:240 */
    TB200_F_D_ALL_HT         = 34,   /* features.py:246 */
    TB200_F_D_MEAN_HT        = 35,   /* features.py:252 */
    TB200_F_D_PITS           = 36,   /* features.py:258 */
    TB200_F_LANDING_HEIGHT   = 37,   /* features.py:266 */
    TB200_F_PATTERN_DIV      = 38,   /* features.py:272 */
    TB200_F_NINE_FILLED      = 39,   /* features.py:291 */
    TB200_F_TETRIS_PROGRESS  = 40,   /* features.py:298 */
    TB200_F_FULL_CELLS       = 41,   /* features.py:324 */
    TB200_F_WEIGHTED_CELLS   = 42,   /* features.py:330 */
    TB200_F_ERODED_CELLS     = 43,   /* features.py:336 */
    TB200_F_CUML_ERODED      = 44,   /* features.py:346 */
    TB200_F__COUNT           = 45
};

#define TB200_FEATURE_NAMES { \
    "mean_ht", "max_ht", "min_ht", "all_ht", "max_ht_diff", \
    "min_ht_diff", "cleared", "cuml_cleared", "tetris", "col_trans", \
    "row_trans", "all_trans", "wells", "cuml_wells", "deep_wells", \
    "max_well", "pits", "pit_rows", "lumped_pits", "pit_depth", \
    "mean_pit_depth", "cd_1", "cd_2", "cd_3", "cd_4", \
    "cd_5", "cd_6", "cd_7", "cd_8", "cd_9", \
    "all_diffs", "max_diffs", "jaggedness", "d_max_ht", "d_all_ht", \
    "d_mean_ht", "d_pits", "landing_height", "pattern_div", "nine_filled", \
    "tetris_progress", "full_cells", "weighted_cells", "eroded_cells", "cuml_eroded" \
}

/* the six-feature "w6" set of the CE configs, in its summation order */
#define TB200_W6_FEATURE_IDS { TB200_F_LANDING_HEIGHT, TB200_F_ERODED_CELLS, TB200_F_ROW_TRANS, TB200_F_COL_TRANS, TB200_F_PITS, TB200_F_CUML_WELLS }

#endif /* TETRIS_B200_FEATURES_H */
