/* alphafive.h -- C ABI of libalphafive.so: the B200-native self-play hot path of
 * GuoYi0/alphaFive (Gomoku MCTS + policy/value net leaf evaluation).
 *
 * The reference is pure Python and has no FFI: its seams are the `pv_fn` callable,
 * the Pipe protocol and the `Player` object.  This header is what a
 * ctypes binding of those seams binds to; every entry point cites the reference
 * code it replaces.  Conventions:
 *   - plain C types only; every function returns 0 on success or a negative
 *     a5_status; a5_last_error() gives the message of the calling thread's last
 *     failure.  Nothing throws, nothing falls back to the CPU.
 *   - pointers named d_* are DEVICE pointers owned by the caller; h_* are host
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - boards are int8[S*S] row-major, +1 = side to move, -1 = opponent, 0 = empty
 *     (utils.py:178-196); a move / last_action is the flat cell i*S+j, -1 = None.
 *   - one host thread per engine; engines are independent (one per GPU / rank).
 */
#ifndef ALPHAFIVE_H_
#define ALPHAFIVE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A5_VERSION 1
#define A5_MAX_BOARD 15          /* S <= 15 (cells <= 256 per warp pass)          */

typedef enum {
  A5_OK = 0,
  A5_ERR_ARG = -1,               /* bad argument                                  */
  A5_ERR_CUDA = -2,              /* CUDA runtime error (see a5_last_error)        */
  A5_ERR_STATE = -3,             /* call not valid in the engine's current mode   */
  A5_ERR_CAPACITY = -4           /* a per-game node arena or the record arena overflowed */
} a5_status;

int a5_version(void);
const char* a5_last_error(void);

/* ---------------------------------------------------------------------------
 * Rules and encodings, batched: one warp per board, board staged in shared memory.
 * ------------------------------------------------------------------------- */

/* utils.py:199-235 is_game_over.  codes[i]: 0 = (False, 0.0), 1 = (True, +1.0),
 * 2 = (True, -1.0), 3 = (True, 0.0) draw.  Scan order as the reference. */
int a5_rules_terminal(const int8_t* d_boards, int n, int S, int goal, int8_t* d_codes, void* stream);

/* utils.py:275-283 step: out = -(board with +1 placed at cell).  d_out may alias d_boards. */
int a5_rules_step(const int8_t* d_boards, const int32_t* d_cells, int n, int S, int8_t* d_out, void* stream);

/* utils.py:238-245 get_legal_actions: mask[i][c] = 1 for empty cells (row-major order
 * is the reference's action order), count[i] = number of legal actions. */
int a5_rules_legal(const int8_t* d_boards, int n, int S, uint8_t* d_mask, int32_t* d_count, void* stream);

/* utils.py:256-272 board_to_inputs: int8[n][3][S*S] planes (own, opponent, one-hot
 * last move); values are exactly {0,1}, so int8 carries the reference's f32 input. */
int a5_rules_inputs(const int8_t* d_boards, const int32_t* d_last, int n, int S, int8_t* d_planes, void* stream);

/* utils.py:156-175 board_to_state: NUL-terminated strings, `stride` bytes apart
 * (stride >= S*(S+1)+1), lengths in d_len. */
int a5_rules_encode(const int8_t* d_boards, int n, int S, char* d_states, int stride, int32_t* d_len, void* stream);

/* utils.py:178-196 state_to_board (inverse of the above). */
int a5_rules_decode(const char* d_states, int stride, int n, int S, int8_t* d_boards, void* stream);

/* ---------------------------------------------------------------------------
 * Search engine: N concurrent games, one warp per game, lock-step simulations.
 * Replaces genData/player.py Player (tree, get_action, MCTS_search,
 * select_action_q_and_u, evaluate_and_expand, update_tree, calc_policy, run).
 * ------------------------------------------------------------------------- */
typedef struct a5_engine a5_engine;

typedef struct {
  int32_t board_size;            /* config.board_size                      config.py:2  */
  int32_t goal;                  /* config.goal                            config.py:6  */
  int32_t n_games;               /* concurrent games (Player instances)                */
  int32_t sims;                  /* config.simulation_per_step             config.py:4  */
  int32_t upper_sims;            /* config.upper_simulation_per_step       config.py:5  */
  float   c_puct;                /* config.c_puct                          config.py:16 */
  float   dirichlet_alpha;       /* config.dirichlet_alpha                 config.py:17 */
  double  init_temp;             /* config.init_temp                       config.py:19 */
  double  tau_decay;             /* config.tau_decay_rate                  config.py:12 */
  double  tau_decay_r;           /* config.tau_decay_rate_r                config.py:15 */
  float   gamma;                 /* config.gamma                           config.py:18 */
  int32_t training;              /* Player(training=...)                 player.py:24  */
  int32_t random_a;              /* get_action(random_a=...)             player.py:128 */
  int32_t auto_play;             /* 1: Player.run() state machine on device: moves are
                                    played, finished games recorded and restarted
                                    (player.py:53-82, main.py:82-94);
                                    0: one get_action per a5_engine_set_roots        */
  int32_t node_capacity;         /* table nodes per game (0 = default)               */
  int32_t max_inner;             /* NN-free (terminal) simulations a game may run in
                                    one step before yielding (0 = default: 2 for
                                    batches of >= 256 games, else 16)                */
  uint64_t seed;                 /* Philox key; stream = game_id_base + game index   */
  int64_t game_id_base;          /* global index of local game 0 (rank offset)       */
  int32_t record_capacity;       /* finished-ply records held for harvest (0 = N*S*S) */
  int32_t reserved;
} a5_config;

int a5_engine_create(const a5_config* cfg, a5_engine** out);
int a5_engine_destroy(a5_engine* e);

/* Player.reset() for every game (player.py:48-51): empty tables, tau = init_temp;
 * in auto_play mode every game restarts from the empty board (player.py:37-46). */
int a5_engine_reset(a5_engine* e, void* stream);

/* auto_play = 0 only.  Player.get_action entry (player.py:138-143) for every game
 * with d_active[i] != 0: set root_state / last_action, keep (and garbage-collect)
 * the existing table, compute the simulation budget
 *   num = sims if the root is unseen else min(sims, upper_sims - root.sum_n).
 * d_clear[i] != 0 first performs Player.reset() for game i.  d_active / d_clear may be NULL
 * (all active / none cleared). */
int a5_engine_set_roots(a5_engine* e, const int8_t* d_boards, const int32_t* d_last,
                        const uint8_t* d_active, const uint8_t* d_clear, void* stream);

/* One lock-step pass of the tree kernel (player.py:204-228,230-279,186-202,166-184).
 * For every game whose previous pass ended at an unseen leaf: expand it with
 * d_prob[i] (f32[S*S], softmax over all cells, unmasked: network.py:88) and back up
 * d_value[i]; then run simulations until the next unseen leaf, whose network input is
 * written to a5_engine_planes()[i] with a5_engine_need_eval()[i] = 1.  Pass NULL
 * prob/value on the very first pass after create/reset/set_roots. */
int a5_engine_step(a5_engine* e, const float* d_prob, const float* d_value, void* stream);
/* The same when not every pending leaf was evaluated this pass (a5_evalcache_lookup): games with
 * d_served[i] == 0 keep their pending leaf and do nothing in this pass. */
int a5_engine_step_served(a5_engine* e, const float* d_prob, const float* d_value, const uint8_t* d_served, void* stream);

/* Per-call arguments of the reference API that are engine state here: the `training`
 * attribute / `random_a` argument of get_action (player.py:24,128) and the lazily read
 * config.simulation_per_step / upper_simulation_per_step (choose_best_player.py:25 mutates
 * them at run time).  Take effect from the next kernel launch. */
int a5_engine_set_mode(a5_engine* e, int training, int random_a);
int a5_engine_set_budget(a5_engine* e, int sims, int upper_sims);

int8_t*  a5_engine_planes(a5_engine* e);      /* int8 [N][3][S*S]   utils.py:256-272 */
uint8_t* a5_engine_need_eval(a5_engine* e);   /* uint8[N]                            */
int32_t* a5_engine_sims_left(a5_engine* e);   /* int32[N] remaining budget           */

/* Number of games that still have simulations to run or a leaf awaiting the
 * network (synchronises the stream). */
int a5_engine_busy(a5_engine* e, int32_t* h_busy, void* stream);

/* auto_play = 0: Player.calc_policy (player.py:84-126) for every active game after
 * its budget is spent.  d_policy f32[N][S*S] (all zero where the reference returns
 * None), d_action int32[N] flat cell. */
int a5_engine_finish_move(a5_engine* e, float* d_policy, int32_t* d_action, void* stream);

/* Continuous batching of get_action calls (auto_play = 0).  The searches of a batch end in different
 * passes: the budget rule (player.py:140-143) gives re-used trees fewer simulations (a trained net
 * concentrates the visits on the move that gets played), simulations that end in terminal positions run
 * two to a pass, leaves deferred by the evaluation cache wait a pass, and roots arrive at different
 * times; the reference's players are independent objects and none of them waits for another (each
 * worker of main.py:50-55 calls
 * get_action on its own).  a5_engine_collect_moves closes every search that has ended and has not been
 * collected yet -- Player.calc_policy (player.py:84-126), then utils.step (utils.py:275-283) and
 * utils.is_game_over (utils.py:199-235) of the played position -- into compact rows:
 *   d_count int32[1] searches found (may exceed cap: the rest is returned by the next call),
 *   d_game int32[cap] game index, d_policy f32[cap][S*S], d_action int32[cap] flat cell,
 *   d_next int8[cap][S*S] position after the move (side to move = +1), d_code int8[cap] its terminal code
 *   (a5_rules_terminal).  A collected game is parked (not searched) until
 * a5_engine_submit_roots gives it its next root: Player.get_action entry for game d_game[i] with board
 * d_boards[i] (int8[n][S*S]), d_last[i] (flat cell or -1; NULL = none) and, where d_clear[i] != 0,
 * Player.reset() first (d_clear may be NULL).  All other games are left untouched and keep searching.
 * d_game must not name a game twice. */
int a5_engine_collect_moves(a5_engine* e, int cap, int32_t* d_count, int32_t* d_game, float* d_policy,
                            int32_t* d_action, int8_t* d_next, int8_t* d_code, void* stream);
int a5_engine_submit_roots(a5_engine* e, int n, const int32_t* d_game, const int8_t* d_boards,
                           const int32_t* d_last, const uint8_t* d_clear, void* stream);

/* Root node statistics of every game as dense per-cell arrays (parity tests, GUI):
 * n int32[N][S*S], w f32, p f32, sum_n int32[N]; any pointer may be NULL. */
int a5_engine_root_stats(a5_engine* e, int32_t* d_n, float* d_w, float* d_p, int32_t* d_sum_n, void* stream);

/* The same for the node of an arbitrary position per game (d_boards int8[N][S*S]): what
 * Player.tree[state].a[action].{n,w,p} / .sum_n hold (player.py:9-20,29); sum_n = -1 and zero rows when the
 * position is not in game i's table.  Read-only. */
int a5_engine_node_stats(a5_engine* e, const int8_t* d_boards, int32_t* d_n, float* d_w, float* d_p,
                         int32_t* d_sum_n, void* stream);

/* The current root position of every game (Player.root_state, player.py:30,138) and the move that led to it:
 * d_boards int8[N][S*S], d_last int32[N] (may be NULL).  In auto_play mode this is where each game stands. */
int a5_engine_get_roots(a5_engine* e, int8_t* d_boards, int32_t* d_last, void* stream);

/* Player.tau of every game (player.py:32,108-111): double[N] on the device, decayed by finish_move. */
double* a5_engine_tau(a5_engine* e);

/* n_draws Dirichlet(alpha * 1_A) vectors, A = n_legal, from the sampler the tree pass uses at every node
 * visit (np.random.dirichlet, player.py:240); draw i uses Philox stream i of `seed`.  d_eta f32[n_draws][n_legal].
 * Exists so the distribution of the exploration noise can be tested against Beta(alpha, alpha (A-1)) marginals. */
int a5_dirichlet_sample(uint64_t seed, float alpha, int n_legal, int n_draws, float* d_eta, void* stream);

/* Table contents of one game: up to max_nodes boards int8[max_nodes][S*S] and their
 * sum_n; returns the number of nodes in *h_count (synchronises). */
int a5_engine_table_dump(a5_engine* e, int game, int8_t* d_boards, int32_t* d_sum_n, int max_nodes,
                         int32_t* h_count, void* stream);

/* Finished-game records (auto_play = 1), the reference's replay record
 * (state, policy, last_action, value, weight) of player.py:77-82 at a fixed stride. */
typedef struct {
  int64_t game_id;               /* global game slot                               */
  int32_t game_serial;           /* how many games this slot finished before       */
  int16_t ply;                   /* index within the game                          */
  int16_t game_len;              /* plies in the game                              */
  int32_t last_action;           /* flat cell or -1 (None)                         */
  float   value;                 /* +-1 / 0, alternating (player.py:75-81)         */
  float   weight;                /* construct_weights (utils.py:286-296)           */
  int32_t result;                /* main.py:86-93: 1 black win, -1 white win, 0 draw */
  /* followed by: int8 board[S*S] padded to a multiple of 16, then f32 policy[S*S] */
} a5_record_header;

int a5_record_stride(int S);     /* bytes per ply record                            */
/* Move up to max_records finished-ply records into d_out (device) and reset the
 * engine's record arena; *h_count = number copied, *h_games = games completed.  A game's plies are
 * always contiguous and complete.  Returns A5_ERR_CAPACITY (after copying what fits) if games were lost
 * because the arena was full or max_records was too small; the loss is also counted in counters()[10]. */
int a5_engine_harvest(a5_engine* e, void* d_out, int max_records, int32_t* h_count, int32_t* h_games, void* stream);

/* Counters since create/reset (synchronises):
 * [0] moves played  [1] simulations  [2] network leaf evaluations  [3] terminal leaves
 * [4] selections (node visits)  [5] sum of legal actions at expansion  [6] games finished
 * [7] max nodes in any table  [8] capacity overflows  [9] tree-kernel passes
 * [10] records dropped  [11..15] reserved */
#define A5_NUM_COUNTERS 16
int a5_engine_counters(a5_engine* e, int64_t* h_out, void* stream);

/* ---------------------------------------------------------------------------
 * Replay buffer sampling.  Replaces the gather / augmentation loop of
 * RandomStack.get_data (utils.py:118-146): d_records is a device array of ply records
 * (a5_record_stride(S) bytes each); sample i is record d_idx[i] under np.rot90(k = d_rot[i])
 * followed by np.flip(axis 0) when d_flip[i] != 0, the last move remapped alike
 * (utils.py:129-140), expanded by board_to_inputs (utils.py:256-272).
 * d_boards f32[num][3][S][S], d_weights f32[num], d_values f32[num], d_policies f32[num][S*S]. */
int a5_replay_sample(const void* d_records, int S, const int64_t* d_idx, const uint8_t* d_rot, const uint8_t* d_flip,
                     int num, float* d_boards, float* d_weights, float* d_values, float* d_policies, void* stream);

/* ---------------------------------------------------------------------------
 * Policy/value network forward.  Replaces ResNet.eval (network.py:90-97), i.e. the
 * pv_fn seam (player.py:190-192) and the batch eval inside NetworkAPI
 * (networkAPI.py:67-68).
 * ------------------------------------------------------------------------- */
typedef struct a5_net a5_net;

#define A5_NET_NUM_TENSORS 42
/* Canonical tensor order: a5_net_tensor_name(i) is the TensorFlow variable name of the
 * shipped checkpoints ("bone/conv1/kernel", ...).  Layouts are TensorFlow's: conv kernels
 * HWIO, dense kernels [in][out]. */
const char* a5_net_tensor_name(int i);
int64_t a5_net_tensor_size(int i, int S);

#define A5_NET_FP32 0            /* fp32 CUDA-core path                              */
#define A5_NET_TC   1            /* tcgen05 tensor-core path, fp16 hi/lo split (3 MMA
                                    passes, fp32 accumulate)                         */
#define A5_NET_SMALL 2           /* latency path for n <= A5_NET_SMALL_MAX boards: the
                                    whole forward as one persistent fp32 kernel -- the
                                    leaf evaluation of ONE Player.get_action search
                                    (player.py:128-147, 190-192; GUI.py:124-167)      */
#define A5_NET_SMALL_MAX 8

int a5_net_create(int S, int max_batch, a5_net** out);
int a5_net_destroy(a5_net* net);
/* d_tensors[i]: device pointer to tensor i (fp32), owned by the caller (PyTorch);
 * the library re-packs them into its kernel layouts on `stream`. */
int a5_net_set_weights(a5_net* net, const float* const* d_tensors, void* stream);
/* prob f32[n][S*S] (softmax over all cells), value f32[n] = tanh(x/2). */
int a5_net_forward(a5_net* net, const int8_t* d_planes, int n, float* d_prob, float* d_value,
                   int mode, void* stream);

/* ---------------------------------------------------------------------------
 * Cross-game evaluation cache for lock-step self-play.  evaluate_and_expand asks pv_fn for every
 * unseen position (player.py:186-202); thousands of games in one batch ask for many positions more
 * than once (openings).  The network is a deterministic, batch-invariant function of its input
 * planes, so serving a leaf from a table of earlier results is exact.  Per pass:
 *   a5_evalcache_lookup  hits -> d_prob / d_value rows of the game; misses -> a compact batch of at
 *                        most `cap` boards (a5_evalcache_planes); leaves that do not fit stay pending
 *                        (d_served[i] = 0, see a5_engine_step_served)
 *   a5_net_forward       on a5_evalcache_planes() with n = cap -> a5_evalcache_prob() / _value()
 *   a5_evalcache_commit  compact results -> the games' rows and into the table
 * a5_evalcache_clear must follow every a5_net_set_weights. */
typedef struct a5_evalcache a5_evalcache;
int a5_evalcache_create(int S, int n_games, int log2_slots, int cap, a5_evalcache** out);
int a5_evalcache_destroy(a5_evalcache* c);
int a5_evalcache_clear(a5_evalcache* c, void* stream);
int a5_evalcache_lookup(a5_evalcache* c, const int8_t* d_planes, const uint8_t* d_need, float* d_prob, float* d_value,
                        uint8_t* d_served, void* stream);
int8_t* a5_evalcache_planes(a5_evalcache* c);
float* a5_evalcache_prob(a5_evalcache* c);
float* a5_evalcache_value(a5_evalcache* c);
int a5_evalcache_commit(a5_evalcache* c, float* d_prob, float* d_value, void* stream);
/* h_out[4]: lookups, hits, leaves deferred (batch full), entries stored -- since create. */
int a5_evalcache_stats(a5_evalcache* c, int64_t* h_out, void* stream);

/* The tensor-core forward in three stream-ordered parts, for callers that pipeline two half batches on
 * SM-partitioned streams (CUDA green contexts): FRONT = input bitboards + conv1, BODY = the nine block
 * convolutions, HEADS = dense heads (writes d_prob / d_value).  a5_net_forward == all three on one stream.
 * a5_net_set_sm_limit tells the persistent kernels how many SMs the stream they will run on owns. */
#define A5_NET_PART_FRONT 1
#define A5_NET_PART_BODY  2
#define A5_NET_PART_HEADS 4
#define A5_NET_PART_ALL   7
int a5_net_forward_parts(a5_net* net, const int8_t* d_planes, int n, float* d_prob, float* d_value, int parts, void* stream);
int a5_net_set_sm_limit(a5_net* net, int body_sms, int front_sms);


#ifdef __cplusplus
}
#endif
#endif /* ALPHAFIVE_H_ */
