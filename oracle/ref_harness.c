/*
 * TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
 *
 * Thin C harness that is linked against the UNMODIFIED reference sources where they
 * lie (the .c files of /root/reference/src minus main.c, see oracle/Makefile) and exposes the three
 * parity functions of SURVEY.md §8(c) plus an annealing-epoch replay through a flat,
 * ctypes-friendly ABI.  Output goes to oracle/_ref/libmegalania_ref.so, which is
 * git-ignored but travels to the GPU box.
 *
 * Nothing here re-implements reference behaviour: every function only drives the
 * reference's own public functions in the order its main.c does
 *   cost  : main.c:116-118 with perplexity_encoder instead of range_encoder
 *   bytes : main.c:110-119
 *   topk  : main.c:53-58 (the commented "demostate" recipe) / neighbour.c:59-70
 *   epoch : main.c:71-102
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lzma_header_encoder.h"
#include "lzma_packet_encoder.h"
#include "lzma_state.h"
#include "max_heap.h"
#include "packet_enumerator.h"
#include "packet_slab.h"
#include "packet_slab_neighbour.h"
#include "packet_slab_undo_stack.h"
#include "perplexity_encoder.h"
#include "range_encoder.h"
#include "substring_enumerator.h"
#include "top_k_packet_finder.h"
#include "perplexity_table.h"

/* ---- byte sink used for the range-coded stream --------------------------------- */
typedef struct {
	uint8_t* buf;
	size_t cap;
	size_t len;
} ByteSink;

static bool sink_write(OutputInterface* out, const void* bytes, size_t count)
{
	ByteSink* sink = (ByteSink*)out->private_data;
	if (sink->len + count <= sink->cap) {
		memcpy(sink->buf + sink->len, bytes, count);
	}
	sink->len += count;
	return true;
}

static void fresh_state(LZMAState* st, const uint8_t* data, size_t n)
{
	LZMAProperties props = { .lc = 0, .lp = 0, .pb = 0 };
	lzma_state_init(st, data, n, props);
}

/* ---- parity function 2: cost(data, slab) ---------------------------------------- */
uint64_t mgref_slab_cost(const uint8_t* data, size_t n, const LZMAPacket* slab)
{
	LZMAState st;
	fresh_state(&st, data, n);
	uint64_t bits = 0;
	EncoderInterface enc;
	perplexity_encoder_new(&enc, &bits);
	while (st.position < st.data_size) {
		lzma_encode_packet(&st, &enc, slab[st.position]);
	}
	return bits;
}

/* cost of the prefix [0, stop) only; stop must be a live packet boundary */
uint64_t mgref_prefix_cost(const uint8_t* data, size_t n, const LZMAPacket* slab, size_t stop)
{
	LZMAState st;
	fresh_state(&st, data, n);
	uint64_t bits = 0;
	EncoderInterface enc;
	perplexity_encoder_new(&enc, &bits);
	while (st.position < stop) {
		lzma_encode_packet(&st, &enc, slab[st.position]);
	}
	return bits;
}

/* ---- parity function 3: bytes(data, slab) --------------------------------------- */
size_t mgref_encode_slab(const uint8_t* data, size_t n, const LZMAPacket* slab, uint8_t* out, size_t cap)
{
	ByteSink sink = { out, cap, 0 };
	OutputInterface output = { sink_write, &sink };
	LZMAState st;
	fresh_state(&st, data, n);
	lzma_encode_header(&st, &output);
	EncoderInterface enc;
	range_encoder_new(&enc, &output);
	while (st.position < st.data_size) {
		lzma_encode_packet(&st, &enc, slab[st.position]);
	}
	range_encoder_free(&enc);
	return sink.len;
}

/* ---- parity function 1: topk(data, state, position, excluded) ------------------- */
/* state_mode 0: freshly initialised model with position forced to pos (main.c:53-57)
 * state_mode 1: model reached by encoding slab[0..pos) with the perplexity back end
 * The exclusion packet is slab[pos], exactly as top_k_packet_finder.c:99 reads it.   */
int mgref_topk(const uint8_t* data, size_t n, const LZMAPacket* slab, int state_mode,
               size_t pos, int k, LZMAPacket* pops)
{
	LZMAState st;
	fresh_state(&st, data, n);
	if (state_mode == 0) {
		st.position = pos;
	} else {
		uint64_t bits = 0;
		EncoderInterface enc;
		perplexity_encoder_new(&enc, &bits);
		while (st.position < pos) {
			lzma_encode_packet(&st, &enc, slab[st.position]);
		}
		if (st.position != pos) {
			return -1;
		}
	}
	PacketEnumerator* en = packet_enumerator_new(data, n);
	TopKPacketFinder* finder = top_k_packet_finder_new((size_t)k, en);
	top_k_packet_finder_find(finder, &st, (LZMAPacket*)slab);
	int count = 0;
	LZMAPacket p;
	while (top_k_packet_finder_pop(finder, &p)) {
		pops[count++] = p;
	}
	top_k_packet_finder_free(finder);
	packet_enumerator_free(en);
	return count;
}

/* Batched form that keeps one enumerator alive (the index build is O(n)). */
int mgref_topk_many(const uint8_t* data, size_t n, const LZMAPacket* slab, int state_mode,
                    const uint64_t* positions, size_t npos, int k, LZMAPacket* pops, int32_t* counts)
{
	PacketEnumerator* en = packet_enumerator_new(data, n);
	TopKPacketFinder* finder = top_k_packet_finder_new((size_t)k, en);
	LZMAState walk;
	fresh_state(&walk, data, n);
	uint64_t bits = 0;
	EncoderInterface enc;
	perplexity_encoder_new(&enc, &bits);
	int rc = 0;
	for (size_t q = 0; q < npos; q++) {
		LZMAState st;
		if (state_mode == 0) {
			fresh_state(&st, data, n);
			st.position = positions[q];
		} else {
			/* positions must be ascending live boundaries */
			while (walk.position < positions[q]) {
				lzma_encode_packet(&walk, &enc, slab[walk.position]);
			}
			if (walk.position != positions[q]) { rc = -1; break; }
			st = walk;
		}
		top_k_packet_finder_find(finder, &st, (LZMAPacket*)slab);
		int c = 0;
		LZMAPacket p;
		while (top_k_packet_finder_pop(finder, &p)) {
			pops[q * (size_t)k + c++] = p;
		}
		counts[q] = c;
	}
	top_k_packet_finder_free(finder);
	packet_enumerator_free(en);
	return rc;
}

/* ---- reference unit-test KATs: substring callback counts ------------------------ */
static void count_cb(void* user, size_t offset, size_t length)
{
	(void)offset; (void)length;
	(*(size_t*)user)++;
}

size_t mgref_substring_count(const uint8_t* data, size_t n, size_t pos, size_t max_len)
{
	SubstringEnumerator* se = substring_enumerator_new(data, n, 2, max_len);
	size_t calls = 0;
	substring_enumerator_for_each(se, pos, count_cb, &calls);
	substring_enumerator_free(se);
	return calls;
}

/* ---- max_heap KAT driver (tests/max_heap_test.c pattern, comparator on plain ints) */
static int int_cmp(void* user, unsigned a, unsigned b)
{
	const int* keys = (const int*)user;
	return (keys[a] > keys[b]) - (keys[a] < keys[b]);
}

/* Streams keys[0..count) through a K-bounded heap using the finder's `<=` replacement
 * rule, then pops; out receives key indices worst-first. */
int mgref_heap_topk(const int* keys, int count, int k, int* out)
{
	int* slot_keys = malloc(sizeof(int) * (size_t)k);
	int* slot_src = malloc(sizeof(int) * (size_t)k);
	MaxHeap* heap = max_heap_new((size_t)k, int_cmp, slot_keys);
	for (int i = 0; i < count; i++) {
		size_t have = max_heap_count(heap);
		if (have < (size_t)k) {
			slot_keys[have] = keys[i];
			slot_src[have] = i;
			max_heap_insert(heap, (unsigned)have);
		} else {
			unsigned top = 0;
			max_heap_maximum(heap, &top);
			if (keys[i] <= slot_keys[top]) {
				slot_keys[top] = keys[i];
				slot_src[top] = i;
				max_heap_update_maximum(heap);
			}
		}
	}
	int produced = 0;
	unsigned top = 0;
	while (max_heap_maximum(heap, &top)) {
		out[produced++] = slot_src[top];
		max_heap_remove_maximum(heap);
	}
	max_heap_free(heap);
	free(slot_keys);
	free(slot_src);
	return produced;
}

/* ---- annealing epoch replay (main.c:71-102) ------------------------------------- */
typedef struct {
	uint64_t cost;       /* neighbour.perplexity (0 when the proposal failed)        */
	uint32_t flags;      /* bit0 generate() returned true, bit1 accepted, bit2 new best */
	uint32_t undo_count; /* packet_slab_neighbour_undo_count before accept/undo       */
} MgTraceRec;

/* slab: current packets (in/out). best/best_cost: in/out like packets_best/best_perplexity.
 * reseed != 0 calls srand(seed) first (main.c:68).  Runs until `evals` proposals succeeded
 * (the loop index i of main.c:78 counts successes) or max_attempts proposals were drawn.
 * num_iters is the value the temperature rule sees (main.c:67, file_size in the stock CLI). */
long mgref_anneal_epoch(const uint8_t* data, size_t n, LZMAPacket* slab, LZMAPacket* best,
                        uint64_t* best_cost, int reseed, unsigned seed, unsigned step,
                        int num_iters, int evals, long max_attempts, uint64_t* cur_cost_io,
                        MgTraceRec* trace, long trace_cap)
{
	LZMAState init_state;
	fresh_state(&init_state, data, n);
	PacketEnumerator* en = packet_enumerator_new(data, n);
	TopKPacketFinder* finder = top_k_packet_finder_new(20, en);
	PacketSlab* ps = packet_slab_new(n);
	LZMAPacket* packets = packet_slab_packets(ps);
	memcpy(packets, slab, sizeof(LZMAPacket) * n);
	if (reseed) {
		srand(seed);
	}
	uint64_t current = *cur_cost_io;
	long attempts = 0;
	int i = 0;
	while (i < evals && attempts < max_attempts) {
		PacketSlabNeighbour nb;
		packet_slab_neighbour_new(&nb, ps, init_state);
		bool ok = packet_slab_neighbour_generate(&nb, finder);
		MgTraceRec rec = { 0, 0, 0 };
		if (ok) {
			bool uphill = rand() % (i*i+1+step*num_iters/2) < sqrt(num_iters);
			rec.cost = nb.perplexity;
			rec.flags = 1;
			rec.undo_count = (uint32_t)packet_slab_neighbour_undo_count(&nb);
			if (current == 0 || nb.perplexity < current || uphill) {
				current = nb.perplexity;
				rec.flags |= 2;
				if (*best_cost == 0 || current < *best_cost) {
					*best_cost = current;
					memcpy(best, packets, sizeof(LZMAPacket) * n);
					rec.flags |= 4;
				}
			} else {
				packet_slab_neighbour_undo(&nb);
			}
			packet_slab_neighbour_free(&nb);
			i++;
		}
		if (attempts < trace_cap && trace != NULL) {
			trace[attempts] = rec;
		}
		attempts++;
	}
	memcpy(slab, packets, sizeof(LZMAPacket) * n);
	*cur_cost_io = current;
	packet_slab_free(ps);
	top_k_packet_finder_free(finder);
	packet_enumerator_free(en);
	return attempts;
}

/* ---- extra pins: price table, full model snapshot, libc rand stream ---------------- */
const uint64_t* mgref_price_table(void) { return LOG2_LOOKUP; }

typedef struct {
	uint16_t probs[sizeof(LZMAProbabilityModel) / sizeof(Prob)];
	uint8_t ctx_state;
	uint32_t dists[4];
	uint64_t position;
	uint64_t cost;
} MgModelDump;

void mgref_model_after_prefix(const uint8_t* data, size_t n, const LZMAPacket* slab, size_t stop,
                              MgModelDump* out)
{
	LZMAState st;
	fresh_state(&st, data, n);
	uint64_t bits = 0;
	EncoderInterface enc;
	perplexity_encoder_new(&enc, &bits);
	while (st.position < stop) {
		lzma_encode_packet(&st, &enc, slab[st.position]);
	}
	memcpy(out->probs, &st.probs, sizeof(out->probs));
	out->ctx_state = st.ctx_state;
	memcpy(out->dists, st.dists, sizeof(st.dists));
	out->position = st.position;
	out->cost = bits;
}

void mgref_rand_stream(unsigned seed, int count, int* out)
{
	srand(seed);
	for (int i = 0; i < count; i++) {
		out[i] = rand();
	}
}

size_t mgref_sizeof_packet(void) { return sizeof(LZMAPacket); }
size_t mgref_sizeof_state(void) { return sizeof(LZMAState); }
