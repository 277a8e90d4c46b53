/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the Megalania annealing hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  The product (megalania_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py) against
 *   - the known answers recorded from the reference (tests/golden/),
 *   - the reference's own unit-test vectors (reference tests/substring_enumerator_test.c:37,59,
 *     tests/max_heap_test.c:93-145),
 *   - the unmodified reference compiled into oracle/_ref/ (when /root/reference is present),
 *   - fixtures generated from that compiled reference (tests/golden/, tools/make_golden.py).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 12-byte layout as reference src/lzma_packet.h:13-17 */
typedef struct {
	uint8_t type;  /* 0 invalid, 1 literal, 2 match, 3 short rep, 4 long rep */
	uint32_t dist; /* match: distance-1; long rep: rep index; else 0 */
	uint16_t len;  /* 1 for literal / short rep, 2..273 otherwise */
} MgoPacket;

typedef struct {
	uint64_t cost;       /* proposal cost (0 when the proposal failed) */
	uint32_t flags;      /* bit0 success, bit1 accepted, bit2 new best */
	uint32_t undo_count; /* number of slab edits the proposal logged   */
} MgoTraceRec;

/* Snapshot of the coder model, probabilities in the reference's struct order
 * (src/lzma_state.h:47-55): lit[768] len[514] rep_len[514] dist[387] ctx[432]. */
#define MGO_NUM_PROBS 2615
typedef struct {
	uint16_t probs[MGO_NUM_PROBS];
	uint8_t ctx_state;
	uint32_t dists[4];
	uint64_t position;
	uint64_t cost;
} MgoModelDump;

/* glibc-compatible rand()/srand() (TYPE_3 additive feedback generator) */
void mgo_srand(unsigned seed);
int mgo_rand(void);

/* The counter-based generator the CUDA chains use (splitmix64, top 31 bits). */
uint64_t mgo_chain_seed(uint64_t seed, uint64_t chain);
uint32_t mgo_chain_rand31(uint64_t* state);

/* floor(-log2(i/2048)*2048), entry 0 = 0: reference generate_table.py:7-9 */
const uint32_t* mgo_price_table(void);

uint64_t mgo_slab_cost(const uint8_t* data, size_t n, const MgoPacket* slab);
uint64_t mgo_prefix_cost(const uint8_t* data, size_t n, const MgoPacket* slab, size_t stop);
void mgo_model_after_prefix(const uint8_t* data, size_t n, const MgoPacket* slab, size_t stop,
                            MgoModelDump* out);
size_t mgo_encode_slab(const uint8_t* data, size_t n, const MgoPacket* slab, uint8_t* out, size_t cap);

/* state_mode 0: fresh model with position forced to pos (reference main.c:53-57);
 * state_mode 1: model reached by pricing slab[0..pos).  Exclusion packet = slab[pos]. */
int mgo_topk(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode, size_t pos,
             int k, MgoPacket* pops);
int mgo_topk_many(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode,
                  const uint64_t* positions, size_t npos, int k, MgoPacket* pops, int32_t* counts);
/* same, also returning the per-byte integer price of every pop */
int mgo_topk_many_priced(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode,
                         const uint64_t* positions, size_t npos, int k, MgoPacket* pops,
                         uint32_t* prices, int32_t* counts);

size_t mgo_substring_count(const uint8_t* data, size_t n, size_t pos, size_t max_len);
int mgo_heap_topk(const int* keys, int count, int k, int* out);

/* One annealing epoch (reference main.c:71-102).
 * rng_mode 0: glibc rand() stream; reseed != 0 seeds it with `seed` first.
 * rng_mode 1: the CUDA chains' generator, state = *rng_state (in/out), reseed/seed ignored.
 * first_eval: index i of the first proposal (the temperature rule reads it).
 * In rng_mode 1 the temperature rule is evaluated in 64-bit integers (no overflow):
 *   uphill = (r % (i*i + 1 + step*num_iters/2))^2 < num_iters                              */
long mgo_anneal_epoch(const uint8_t* data, size_t n, MgoPacket* slab, MgoPacket* best,
                      uint64_t* best_cost, int rng_mode, int reseed, unsigned seed,
                      uint64_t* rng_state, unsigned step, int num_iters, int first_eval, int evals,
                      long max_attempts, uint64_t* cur_cost_io, MgoTraceRec* trace, long trace_cap);

/* Event stream of a slab: the exact sequence of EncoderInterface calls the reference would make
 * (encode_bit(bit, prob) / encode_direct_bits(bits, n)), packed as
 *   modelled bit : (bit << 15) | prob
 *   direct bits  : 0x4000 | nbits, followed by two words holding bits low16 / high16
 * Returns the number of u16 words (written only while < cap). */
size_t mgo_slab_events(const uint8_t* data, size_t n, const MgoPacket* slab, uint16_t* out, size_t cap);

/* Builds a deterministic non-trivial valid slab: at every live position take the cheapest
 * top-k candidate (the last pop) under the running model. */
void mgo_greedy_slab(const uint8_t* data, size_t n, MgoPacket* slab);

size_t mgo_live_count(const MgoPacket* slab, size_t n);

/* 1 if every live packet of the slab is decodable (matches/reps reproduce the data). */
int mgo_slab_valid(const uint8_t* data, size_t n, const MgoPacket* slab);

/* Finder limits beyond the reference (it has none: its window test is commented out, src/substring_enumerator.c:97):
 * window = farthest match start in bytes (0 = unlimited), max_occ = only the nearest max_occ earlier occurrences of
 * the bigram (0 = all).  Process-global, off by default; checks mg_ctx_set_finder_limits. */
void mgo_set_finder_limits(size_t window, uint32_t max_occ);

/* Literal context bits, 0..4 (0 = the reference).  Process-global. */
void mgo_set_lc(unsigned lc);
unsigned mgo_get_lc(void);

#ifdef __cplusplus
}
#endif
