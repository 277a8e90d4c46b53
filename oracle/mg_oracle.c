/*
 * TEST INFRASTRUCTURE ONLY — see mg_oracle.h.  CPU restatement of the Megalania annealing
 * hot path, written from the reference's behaviour (citations per function), not from its
 * text: the model is a flat slot array, the coder back ends are a tagged sink, the finder
 * prices candidates from the pre-state without copying it.
 *
 * Parity status: PINNED (tests/test_oracle_*.py: reference KATs, the compiled reference
 * in oracle/_ref, and fixtures generated from it).
 */
#include "mg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { T_INVALID = 0, T_LITERAL = 1, T_MATCH = 2, T_SHORT_REP = 3, T_LONG_REP = 4 };

/* ---- slot map: reference struct order, src/lzma_state.h:15-55 ------------------------ */
enum {
	LEN_CHOICE1 = 0,
	LEN_CHOICE2 = 1,
	LEN_LOW = 2,              /* [16][8]  */
	LEN_MID = 2 + 128,        /* [16][8]  */
	LEN_HIGH = 2 + 256,       /* [256]    */
	LEN_SLOTS = 2 + 256 + 256 /* 514      */
};
enum {
	S_LIT = 0,
	S_LEN = 768,
	S_REPLEN = S_LEN + LEN_SLOTS,
	S_POSSLOT = S_REPLEN + LEN_SLOTS, /* [4][64] */
	S_ALIGN = S_POSSLOT + 256,        /* [16]    */
	S_POSCODER = S_ALIGN + 16,        /* [115]   */
	S_ISMATCH = S_POSCODER + 115,     /* [12<<4] */
	S_ISREP = S_ISMATCH + 192,
	S_ISREPG0 = S_ISREP + 12,
	S_ISREPG1 = S_ISREPG0 + 12,
	S_ISREPG2 = S_ISREPG1 + 12,
	S_ISREP0LONG = S_ISREPG2 + 12, /* [12<<4] */
	S_TOTAL = S_ISREP0LONG + 192
};
_Static_assert(S_TOTAL == MGO_NUM_PROBS, "slot map must cover the reference model");

/* Literal context bits (lc).  The reference is fixed at lc = lp = pb = 0 (its todos: src/lzma_packet_encoder.c:17,
 * 44,113); with lc > 0 the literal coder of LZMA picks one of 1 << lc tables by the top lc bits of the previous
 * byte.  Table 0 keeps the reference's place, tables 1.. follow the reference's model.  Process-global, 0 by default;
 * pinned, for want of a reference, by decoding the streams with liblzma (tests/test_oracle_golden.py). */
#define MGO_MAX_LC 4
static unsigned g_lc = 0;
void mgo_set_lc(unsigned lc) { g_lc = lc > MGO_MAX_LC ? MGO_MAX_LC : lc; }
unsigned mgo_get_lc(void) { return g_lc; }

typedef struct {
	const uint8_t* data;
	size_t n;
	uint16_t p[S_TOTAL + ((1 << MGO_MAX_LC) - 1) * 768];
	uint8_t ctx;
	uint32_t rep[4];
	size_t pos;
} Model;

/* ---- sinks ---------------------------------------------------------------------------- */
enum { SINK_COST, SINK_RC, SINK_EVENTS };
typedef struct {
	int kind;
	uint64_t cost;
	/* range coder, reference src/range_encoder.c:10-16,83-95 */
	uint64_t low;
	uint32_t range;
	uint8_t cache;
	uint64_t cache_size;
	uint8_t* out;
	size_t out_cap, out_len;
	/* event recorder */
	uint16_t* ev;
	size_t ev_cap, ev_len;
} Sink;

static uint32_t g_price[2048];
static int g_price_ready = 0;

const uint32_t* mgo_price_table(void)
{
	if (!g_price_ready) {
		/* generate_table.py:7-9: -int(log2(i/2048.)*2048), int() truncates toward zero */
		g_price[0] = 0;
		for (int i = 1; i < 2048; i++) {
			double v = log2((double)i / 2048.0) * 2048.0;
			g_price[i] = (uint32_t)(-(long)v);
		}
		g_price_ready = 1;
	}
	return g_price;
}

static void sink_init(Sink* s, int kind)
{
	memset(s, 0, sizeof(*s));
	s->kind = kind;
	s->range = 0xFFFFFFFFu;
	s->cache_size = 1;
	mgo_price_table();
}

static void rc_put(Sink* s, uint8_t b)
{
	if (s->out_len < s->out_cap) s->out[s->out_len] = b;
	s->out_len++;
}

/* src/range_encoder.c:18-38 */
static void rc_shift_low(Sink* s)
{
	uint32_t hi = (uint32_t)(s->low >> 32);
	uint32_t lo = (uint32_t)s->low;
	if (lo < 0xFF000000u || hi != 0) {
		uint8_t carry_base = s->cache;
		do {
			rc_put(s, (uint8_t)(carry_base + (hi & 0xFF)));
			carry_base = 0xFF;
		} while (--s->cache_size != 0);
		s->cache = (uint8_t)(s->low >> 24);
	}
	s->cache_size++;
	s->low = (uint64_t)(lo << 8);
}

static void ev_put(Sink* s, uint16_t w)
{
	if (s->ev_len < s->ev_cap) s->ev[s->ev_len] = w;
	s->ev_len++;
}

/* perplexity_encoder.c:6-10 / range_encoder.c:47-64 */
static void sink_bit(Sink* s, unsigned bit, unsigned prob)
{
	switch (s->kind) {
	case SINK_COST:
		s->cost += g_price[bit ? 2048 - prob : prob];
		break;
	case SINK_RC: {
		uint32_t bound = (s->range >> 11) * prob;
		if (bit) {
			s->low += bound;
			s->range -= bound;
		} else {
			s->range = bound;
		}
		while ((s->range & 0xFF000000u) == 0) {
			s->range <<= 8;
			rc_shift_low(s);
		}
		break;
	}
	case SINK_EVENTS:
		ev_put(s, (uint16_t)((bit << 15) | prob));
		break;
	}
}

/* perplexity_encoder.c:12-17 / range_encoder.c:66-81 */
static void sink_direct(Sink* s, unsigned bits, unsigned nbits)
{
	switch (s->kind) {
	case SINK_COST:
		s->cost += (uint64_t)nbits << 11;
		break;
	case SINK_RC:
		do {
			unsigned bit = (bits >> (nbits - 1)) & 1;
			s->range >>= 1;
			if (bit) s->low += s->range;
			if ((s->range & 0xFF000000u) == 0) {
				s->range <<= 8;
				rc_shift_low(s);
			}
		} while (--nbits);
		break;
	case SINK_EVENTS:
		ev_put(s, (uint16_t)(0x4000 | nbits));
		ev_put(s, (uint16_t)(bits & 0xFFFF));
		ev_put(s, (uint16_t)(bits >> 16));
		break;
	}
}

/* ---- model ---------------------------------------------------------------------------- */
static void model_init(Model* m, const uint8_t* data, size_t n)
{
	m->data = data;
	m->n = n;
	for (size_t i = 0; i < sizeof(m->p) / sizeof(m->p[0]); i++) m->p[i] = 1024; /* probability.h:7 */
	m->ctx = 0;
	memset(m->rep, 0, sizeof(m->rep));
	m->pos = 0;
}

/* first slot of the literal table the packet at m->pos uses: the top lc bits of the previous byte (0 before byte 0) */
static unsigned literal_base(const Model* m)
{
	if (g_lc == 0 || m->pos == 0) return S_LIT;
	unsigned t = m->data[m->pos - 1] >> (8 - g_lc);
	return t == 0 ? S_LIT : S_TOTAL + (t - 1) * 768;
}

/* probability_model.c:5-15 */
static void code_bit(Model* m, Sink* s, unsigned slot, unsigned bit)
{
	unsigned v = m->p[slot];
	sink_bit(s, bit, v);
	if (bit) v -= v >> 5;
	else v += (2048 - v) >> 5;
	m->p[slot] = (uint16_t)v;
}

/* probability_model.c:22-32 */
static void code_tree(Model* m, Sink* s, unsigned base, unsigned nbits, unsigned value)
{
	unsigned node = 1;
	for (unsigned i = nbits; i-- > 0;) {
		unsigned bit = (value >> i) & 1;
		code_bit(m, s, base + node, bit);
		node = (node << 1) | bit;
	}
}

/* probability_model.c:34-44 */
static void code_tree_rev(Model* m, Sink* s, unsigned base, unsigned nbits, unsigned value)
{
	unsigned node = 1;
	for (unsigned i = 0; i < nbits; i++) {
		unsigned bit = value & 1;
		code_bit(m, s, base + node, bit);
		node = (node << 1) | bit;
		value >>= 1;
	}
}

/* lzma_packet_encoder.c:42-63 (pos_state fixed to 0) */
static void code_length(Model* m, Sink* s, unsigned base, unsigned len)
{
	unsigned v = len - 2;
	if (v < 8) {
		code_bit(m, s, base + LEN_CHOICE1, 0);
		code_tree(m, s, base + LEN_LOW, 3, v);
		return;
	}
	v -= 8;
	code_bit(m, s, base + LEN_CHOICE1, 1);
	if (v < 8) {
		code_bit(m, s, base + LEN_CHOICE2, 0);
		code_tree(m, s, base + LEN_MID, 3, v);
		return;
	}
	v -= 8;
	code_bit(m, s, base + LEN_CHOICE2, 1);
	code_tree(m, s, base + LEN_HIGH, 8, v);
}

static unsigned bit_length(unsigned v) { return 32u - (unsigned)__builtin_clz(v); }

/* lzma_packet_encoder.c:71-104 */
static void code_distance(Model* m, Sink* s, unsigned dist, unsigned len)
{
	unsigned lctx = len - 2;
	if (lctx > 3) lctx = 3;
	unsigned slot_base = S_POSSLOT + lctx * 64;
	if (dist < 4) {
		code_tree(m, s, slot_base, 6, dist);
		return;
	}
	unsigned nlow = bit_length(dist) - 2;
	unsigned low = dist & ((1u << nlow) - 1);
	unsigned high = dist >> nlow;
	unsigned pslot = nlow * 2 + high;
	code_tree(m, s, slot_base, 6, pslot);
	if (pslot < 14) {
		unsigned off = (high << nlow) - pslot;
		code_tree_rev(m, s, S_POSCODER + off, nlow, low);
		return;
	}
	sink_direct(s, low >> 4, nlow - 4);
	code_tree_rev(m, s, S_ALIGN, 4, low & 15);
}

/* lzma_state.c:29-57 */
static unsigned next_ctx(unsigned ctx, unsigned type)
{
	switch (type) {
	case T_LITERAL: return ctx < 4 ? 0 : (ctx < 10 ? ctx - 3 : ctx - 6);
	case T_MATCH: return ctx < 7 ? 7 : 10;
	case T_SHORT_REP: return ctx < 7 ? 9 : 11;
	default: return ctx < 7 ? 8 : 11;
	}
}

/* lzma_packet_encoder.c:106-194 */
static void code_packet(Model* m, Sink* s, MgoPacket pk)
{
	unsigned ctx = m->ctx;
	switch (pk.type) {
	case T_LITERAL: {
		code_bit(m, s, S_ISMATCH + (ctx << 4), 0);
		unsigned byte = m->data[m->pos];
		int matched = ctx >= 7;
		unsigned mbyte = matched ? m->data[m->pos - m->rep[0] - 1] : 0;
		unsigned node = 1;
		for (int i = 7; i >= 0; i--) {
			unsigned bit = (byte >> i) & 1;
			unsigned slot = node;
			if (matched) {
				unsigned mbit = (mbyte >> i) & 1;
				slot += (1 + mbit) << 8;
				matched = (mbit == bit);
			}
			code_bit(m, s, literal_base(m) + slot, bit);
			node = (node << 1) | bit;
		}
		break;
	}
	case T_MATCH:
		code_bit(m, s, S_ISMATCH + (ctx << 4), 1);
		code_bit(m, s, S_ISREP + ctx, 0);
		m->rep[3] = m->rep[2]; /* lzma_state.c:59-65 */
		m->rep[2] = m->rep[1];
		m->rep[1] = m->rep[0];
		m->rep[0] = pk.dist;
		code_length(m, s, S_LEN, pk.len);
		code_distance(m, s, pk.dist, pk.len);
		break;
	case T_SHORT_REP:
		code_bit(m, s, S_ISMATCH + (ctx << 4), 1);
		code_bit(m, s, S_ISREP + ctx, 1);
		code_bit(m, s, S_ISREPG0 + ctx, 0);
		code_bit(m, s, S_ISREP0LONG + (ctx << 4), 0);
		break;
	case T_LONG_REP: {
		unsigned idx = pk.dist;
		code_bit(m, s, S_ISMATCH + (ctx << 4), 1);
		code_bit(m, s, S_ISREP + ctx, 1);
		if (idx == 0) {
			code_bit(m, s, S_ISREPG0 + ctx, 0);
			code_bit(m, s, S_ISREP0LONG + (ctx << 4), 1);
		} else {
			code_bit(m, s, S_ISREPG0 + ctx, 1);
			code_bit(m, s, S_ISREPG1 + ctx, idx != 1);
			if (idx != 1) code_bit(m, s, S_ISREPG2 + ctx, idx != 2);
		}
		uint32_t d = m->rep[idx]; /* lzma_state.c:67-81 */
		for (unsigned i = idx; i > 0; i--) m->rep[i] = m->rep[i - 1];
		m->rep[0] = d;
		code_length(m, s, S_REPLEN, pk.len);
		break;
	}
	default:
		abort();
	}
	m->ctx = (uint8_t)next_ctx(ctx, pk.type);
	m->pos += pk.len;
}

/* ---- parity function 2: cost ---------------------------------------------------------- */
uint64_t mgo_prefix_cost(const uint8_t* data, size_t n, const MgoPacket* slab, size_t stop)
{
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_COST);
	while (m.pos < stop) code_packet(&m, &s, slab[m.pos]);
	return s.cost;
}

uint64_t mgo_slab_cost(const uint8_t* data, size_t n, const MgoPacket* slab)
{
	return mgo_prefix_cost(data, n, slab, n);
}

void mgo_model_after_prefix(const uint8_t* data, size_t n, const MgoPacket* slab, size_t stop,
                            MgoModelDump* out)
{
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_COST);
	while (m.pos < stop) code_packet(&m, &s, slab[m.pos]);
	memcpy(out->probs, m.p, sizeof(out->probs)); /* the reference-layout part (lc = 0 tables) */
	out->ctx_state = m.ctx;
	memcpy(out->dists, m.rep, sizeof(m.rep));
	out->position = m.pos;
	out->cost = s.cost;
}

size_t mgo_slab_events(const uint8_t* data, size_t n, const MgoPacket* slab, uint16_t* out, size_t cap)
{
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_EVENTS);
	s.ev = out;
	s.ev_cap = cap;
	while (m.pos < n) code_packet(&m, &s, slab[m.pos]);
	return s.ev_len;
}

/* ---- parity function 3: bytes (main.c:110-119, lzma_header_encoder.c:5-21) ------------- */
size_t mgo_encode_slab(const uint8_t* data, size_t n, const MgoPacket* slab, uint8_t* out, size_t cap)
{
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_RC);
	s.out = out;
	s.out_cap = cap;
	rc_put(&s, (uint8_t)g_lc); /* props: (pb*5+lp)*9+lc with lp=pb=0 */
	uint32_t dict = 0x400000;
	for (int i = 0; i < 4; i++) rc_put(&s, (uint8_t)(dict >> (8 * i)));
	/* the reference writes htole32(size) widened to 8 bytes: high word is always zero */
	uint64_t sz = (uint32_t)n;
	for (int i = 0; i < 8; i++) rc_put(&s, (uint8_t)(sz >> (8 * i)));
	while (m.pos < n) code_packet(&m, &s, slab[m.pos]);
	for (int i = 0; i < 5; i++) rc_shift_low(&s); /* range_encoder.c:40-45 */
	return s.out_len;
}

/* ---- bigram index (substring_enumerator.c:26-47) -------------------------------------- */
typedef struct {
	const uint8_t* data;
	size_t n;
	uint32_t* start; /* 65536+1, keyed by first<<8|second */
	uint32_t* occ;   /* positions, ascending inside a bucket */
} Index;

static void index_build(Index* ix, const uint8_t* data, size_t n)
{
	ix->data = data;
	ix->n = n;
	ix->start = calloc(65537, sizeof(uint32_t));
	ix->occ = malloc(sizeof(uint32_t) * (n ? n : 1));
	for (size_t i = 0; i + 1 < n; i++) ix->start[((unsigned)data[i] << 8 | data[i + 1]) + 1]++;
	for (int b = 0; b < 65536; b++) ix->start[b + 1] += ix->start[b];
	uint32_t* fill = malloc(sizeof(uint32_t) * 65536);
	memcpy(fill, ix->start, sizeof(uint32_t) * 65536);
	for (size_t i = 0; i + 1 < n; i++) ix->occ[fill[(unsigned)data[i] << 8 | data[i + 1]]++] = (uint32_t)i;
	free(fill);
}

static void index_free(Index* ix)
{
	free(ix->start);
	free(ix->occ);
}

size_t mgo_substring_count(const uint8_t* data, size_t n, size_t pos, size_t max_len)
{
	/* substring_enumerator.c:85-105 */
	Index ix;
	index_build(&ix, data, n);
	size_t calls = 0;
	if (pos != 0 && pos != n - 1 && pos < n) {
		unsigned key = (unsigned)data[pos] << 8 | data[pos + 1];
		for (uint32_t i = ix.start[key]; i < ix.start[key + 1]; i++) {
			size_t o = ix.occ[i];
			if (o >= pos) break;
			calls++;
			for (size_t j = 2; j < max_len && j + pos < n; j++) {
				if (data[pos + j] != data[o + j]) break;
				calls++;
			}
		}
	}
	index_free(&ix);
	return calls;
}

/* ---- bounded max-heap with the reference's tie behaviour (max_heap.c:82-167) ----------- */
typedef struct {
	MgoPacket packet;
	uint32_t price; /* cost / len, integer; reference keeps it in a float, exact below 2^24 */
} Entry;

typedef struct {
	int k, count;
	Entry* entries;
	unsigned* store;
} TopK;

static void topk_init(TopK* t, int k)
{
	t->k = k;
	t->count = 0;
	t->entries = malloc(sizeof(Entry) * (size_t)k);
	t->store = malloc(sizeof(unsigned) * (size_t)k);
}

static void topk_free(TopK* t)
{
	free(t->entries);
	free(t->store);
}

static int heap_gt(const TopK* t, unsigned a, unsigned b)
{
	return t->entries[a].price > t->entries[b].price;
}

static void heap_sift_down(TopK* t, size_t parent)
{
	for (;;) {
		size_t l = parent * 2 + 1, r = l + 1;
		if (l >= (size_t)t->count) break;
		size_t big = l;
		if (r < (size_t)t->count && heap_gt(t, t->store[r], t->store[l])) big = r;
		if (!heap_gt(t, t->store[big], t->store[parent])) break;
		unsigned tmp = t->store[big];
		t->store[big] = t->store[parent];
		t->store[parent] = tmp;
		parent = big;
	}
}

static void heap_sift_up(TopK* t, size_t node)
{
	while (node > 0) {
		size_t parent = (node - 1) / 2;
		if (!heap_gt(t, t->store[node], t->store[parent])) break;
		unsigned tmp = t->store[node];
		t->store[node] = t->store[parent];
		t->store[parent] = tmp;
		node = parent;
	}
}

/* top_k_packet_finder.c:72-93 */
static void topk_offer(TopK* t, MgoPacket pk, uint32_t price)
{
	if (t->count < t->k) {
		int slot = t->count;
		t->entries[slot].packet = pk;
		t->entries[slot].price = price;
		t->store[t->count++] = (unsigned)slot;
		heap_sift_up(t, (size_t)slot);
		return;
	}
	unsigned top = t->store[0];
	if (price <= t->entries[top].price) {
		t->entries[top].packet = pk;
		t->entries[top].price = price;
		heap_sift_down(t, 0);
	}
}

/* top_k_packet_finder.c:127-138 */
static int topk_pop(TopK* t, MgoPacket* pk, uint32_t* price)
{
	if (t->count == 0) return 0;
	unsigned top = t->store[0];
	*pk = t->entries[top].packet;
	if (price) *price = t->entries[top].price;
	t->store[0] = t->store[--t->count];
	heap_sift_down(t, 0);
	return 1;
}

int mgo_heap_topk(const int* keys, int count, int k, int* out)
{
	TopK t;
	topk_init(&t, k);
	int* src = malloc(sizeof(int) * (size_t)k);
	for (int i = 0; i < count; i++) {
		MgoPacket dummy = { 0, 0, 0 };
		int before = t.count;
		unsigned top = before ? t.store[0] : 0;
		uint32_t price = (uint32_t)(keys[i] + 0x40000000);
		if (before < k) {
			src[before] = i;
		} else if (price <= t.entries[top].price) {
			src[top] = i;
		}
		topk_offer(&t, dummy, price);
	}
	int produced = 0;
	while (t.count) {
		out[produced++] = src[t.store[0]];
		MgoPacket pk;
		topk_pop(&t, &pk, NULL);
	}
	free(src);
	topk_free(&t);
	return produced;
}

/* ---- candidate pricing without copying the model -------------------------------------- */
/* Every slot is touched at most once per packet, so a packet's price is a pure function of
 * the pre-state (lzma_packet_encoder.c:21-38,48-61,80-103,123-135). */
static uint32_t bit_price(const Model* m, unsigned slot, unsigned bit)
{
	unsigned v = m->p[slot];
	return g_price[bit ? 2048 - v : v];
}

static uint32_t tree_price(const Model* m, unsigned base, unsigned nbits, unsigned value)
{
	uint32_t c = 0;
	unsigned node = 1;
	for (unsigned i = nbits; i-- > 0;) {
		unsigned bit = (value >> i) & 1;
		c += bit_price(m, base + node, bit);
		node = (node << 1) | bit;
	}
	return c;
}

static uint32_t tree_rev_price(const Model* m, unsigned base, unsigned nbits, unsigned value)
{
	uint32_t c = 0;
	unsigned node = 1;
	for (unsigned i = 0; i < nbits; i++) {
		unsigned bit = value & 1;
		c += bit_price(m, base + node, bit);
		node = (node << 1) | bit;
		value >>= 1;
	}
	return c;
}

static uint32_t length_price(const Model* m, unsigned base, unsigned len)
{
	unsigned v = len - 2;
	if (v < 8) return bit_price(m, base + LEN_CHOICE1, 0) + tree_price(m, base + LEN_LOW, 3, v);
	v -= 8;
	uint32_t c = bit_price(m, base + LEN_CHOICE1, 1);
	if (v < 8) return c + bit_price(m, base + LEN_CHOICE2, 0) + tree_price(m, base + LEN_MID, 3, v);
	v -= 8;
	return c + bit_price(m, base + LEN_CHOICE2, 1) + tree_price(m, base + LEN_HIGH, 8, v);
}

static uint32_t distance_price(const Model* m, unsigned dist, unsigned len)
{
	unsigned lctx = len - 2;
	if (lctx > 3) lctx = 3;
	unsigned slot_base = S_POSSLOT + lctx * 64;
	if (dist < 4) return tree_price(m, slot_base, 6, dist);
	unsigned nlow = bit_length(dist) - 2;
	unsigned low = dist & ((1u << nlow) - 1);
	unsigned high = dist >> nlow;
	unsigned pslot = nlow * 2 + high;
	uint32_t c = tree_price(m, slot_base, 6, pslot);
	if (pslot < 14) return c + tree_rev_price(m, S_POSCODER + (high << nlow) - pslot, nlow, low);
	return c + ((nlow - 4) << 11) + tree_rev_price(m, S_ALIGN, 4, low & 15);
}

static uint32_t literal_price(const Model* m)
{
	unsigned ctx = m->ctx;
	uint32_t c = bit_price(m, S_ISMATCH + (ctx << 4), 0);
	unsigned byte = m->data[m->pos];
	int matched = ctx >= 7;
	unsigned mbyte = matched ? m->data[m->pos - m->rep[0] - 1] : 0;
	unsigned node = 1;
	for (int i = 7; i >= 0; i--) {
		unsigned bit = (byte >> i) & 1;
		unsigned slot = node;
		if (matched) {
			unsigned mbit = (mbyte >> i) & 1;
			slot += (1 + mbit) << 8;
			matched = (mbit == bit);
		}
		c += bit_price(m, literal_base(m) + slot, bit);
		node = (node << 1) | bit;
	}
	return c;
}

static uint32_t packet_price(const Model* m, MgoPacket pk)
{
	unsigned ctx = m->ctx;
	switch (pk.type) {
	case T_LITERAL: return literal_price(m);
	case T_MATCH:
		return bit_price(m, S_ISMATCH + (ctx << 4), 1) + bit_price(m, S_ISREP + ctx, 0) +
		       length_price(m, S_LEN, pk.len) + distance_price(m, pk.dist, pk.len);
	case T_SHORT_REP:
		return bit_price(m, S_ISMATCH + (ctx << 4), 1) + bit_price(m, S_ISREP + ctx, 1) +
		       bit_price(m, S_ISREPG0 + ctx, 0) + bit_price(m, S_ISREP0LONG + (ctx << 4), 0);
	default: {
		unsigned idx = pk.dist;
		uint32_t c = bit_price(m, S_ISMATCH + (ctx << 4), 1) + bit_price(m, S_ISREP + ctx, 1);
		if (idx == 0) {
			c += bit_price(m, S_ISREPG0 + ctx, 0) + bit_price(m, S_ISREP0LONG + (ctx << 4), 1);
		} else {
			c += bit_price(m, S_ISREPG0 + ctx, 1) + bit_price(m, S_ISREPG1 + ctx, idx != 1);
			if (idx != 1) c += bit_price(m, S_ISREPG2 + ctx, idx != 2);
		}
		return c + length_price(m, S_REPLEN, pk.len);
	}
	}
}

static int packet_eq(MgoPacket a, MgoPacket b)
{
	return a.type == b.type && a.len == b.len && a.dist == b.dist;
}

static MgoPacket mk(unsigned type, unsigned dist, unsigned len)
{
	MgoPacket p;
	memset(&p, 0, sizeof(p));
	p.type = (uint8_t)type;
	p.dist = dist;
	p.len = (uint16_t)len;
	return p;
}

/* Finder limits for inputs beyond the reference's reach: 0 = none = the reference's semantics. */
static size_t g_finder_window = 0;
static uint32_t g_finder_max_occ = 0;
void mgo_set_finder_limits(size_t window, uint32_t max_occ)
{
	g_finder_window = window;
	g_finder_max_occ = max_occ;
}

/* top_k_packet_finder.c:95-118 */
static void consider(TopK* t, const Model* m, MgoPacket excluded, MgoPacket pk)
{
	if (packet_eq(pk, excluded)) return;
	uint32_t price = packet_price(m, pk) / pk.len;
	topk_offer(t, pk, price);
}

/* packet_enumerator.c:41-74 + substring_enumerator.c:85-105 + top_k_packet_finder.c:120-125 */
static void find_candidates(TopK* t, const Index* ix, const Model* m, MgoPacket excluded)
{
	const uint8_t* data = m->data;
	size_t n = m->n, pos = m->pos;
	t->count = 0;
	consider(t, m, excluded, mk(T_LITERAL, 0, 1));
	if (pos > 0 && data[pos] == data[pos - m->rep[0] - 1]) consider(t, m, excluded, mk(T_SHORT_REP, 0, 1));
	if (pos == 0 || pos == n - 1) return;
	unsigned key = (unsigned)data[pos] << 8 | data[pos + 1];
	uint32_t first = ix->start[key];
	if (g_finder_max_occ != 0) {
		/* only the nearest max_occ earlier occurrences (an extension beyond the reference, off by default) */
		uint32_t upper = first;
		while (upper < ix->start[key + 1] && ix->occ[upper] < pos) upper++;
		if (upper - first > g_finder_max_occ) first = upper - g_finder_max_occ;
	}
	for (uint32_t i = first; i < ix->start[key + 1]; i++) {
		size_t o = ix->occ[i];
		if (o >= pos) break;
		/* the window limit the reference leaves commented out (src/substring_enumerator.c:97), off by default */
		if (g_finder_window != 0 && pos - o > g_finder_window) continue;
		unsigned dist = (unsigned)(pos - o - 1);
		for (size_t len = 2; len <= 273 && pos + len <= n; len++) {
			if (len > 2 && data[pos + len - 1] != data[o + len - 1]) break;
			consider(t, m, excluded, mk(T_MATCH, dist, (unsigned)len));
			for (unsigned r = 0; r < 4; r++)
				if (m->rep[r] == dist) consider(t, m, excluded, mk(T_LONG_REP, r, (unsigned)len));
		}
	}
}

static int topk_run(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode,
                    const uint64_t* positions, size_t npos, int k, MgoPacket* pops, uint32_t* prices,
                    int32_t* counts)
{
	Index ix;
	index_build(&ix, data, n);
	TopK t;
	topk_init(&t, k);
	Model walk;
	Sink s;
	model_init(&walk, data, n);
	sink_init(&s, SINK_COST);
	int rc = 0;
	Model* fresh = malloc(sizeof(Model));
	for (size_t q = 0; q < npos; q++) {
		const Model* m;
		if (state_mode == 0) {
			model_init(fresh, data, n);
			fresh->pos = positions[q];
			m = fresh;
		} else {
			while (walk.pos < positions[q]) code_packet(&walk, &s, slab[walk.pos]);
			if (walk.pos != positions[q]) {
				rc = -1;
				break;
			}
			m = &walk;
		}
		find_candidates(&t, &ix, m, slab[m->pos]);
		int c = 0;
		MgoPacket pk;
		uint32_t price;
		while (topk_pop(&t, &pk, &price)) {
			pops[q * (size_t)k + c] = pk;
			if (prices) prices[q * (size_t)k + c] = price;
			c++;
		}
		counts[q] = c;
	}
	free(fresh);
	topk_free(&t);
	index_free(&ix);
	return rc;
}

int mgo_topk_many(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode,
                  const uint64_t* positions, size_t npos, int k, MgoPacket* pops, int32_t* counts)
{
	return topk_run(data, n, slab, state_mode, positions, npos, k, pops, NULL, counts);
}

int mgo_topk_many_priced(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode,
                         const uint64_t* positions, size_t npos, int k, MgoPacket* pops,
                         uint32_t* prices, int32_t* counts)
{
	return topk_run(data, n, slab, state_mode, positions, npos, k, pops, prices, counts);
}

int mgo_topk(const uint8_t* data, size_t n, const MgoPacket* slab, int state_mode, size_t pos, int k,
             MgoPacket* pops)
{
	uint64_t p = pos;
	int32_t count = 0;
	if (topk_run(data, n, slab, state_mode, &p, 1, k, pops, NULL, &count) < 0) return -1;
	return count;
}

/* ---- RNGs ------------------------------------------------------------------------------ */
/* glibc random_r TYPE_3: r[i] = r[i-3] + r[i-31], 310 outputs discarded, result >> 1 */
static int32_t g_r[34];
static int g_f = 3, g_b = 0;

void mgo_srand(unsigned seed)
{
	if (seed == 0) seed = 1;
	int32_t word = (int32_t)seed;
	g_r[0] = word;
	for (int i = 1; i < 31; i++) {
		long hi = word / 127773, lo = word % 127773;
		long w = 16807 * lo - 2836 * hi;
		if (w < 0) w += 2147483647;
		word = (int32_t)w;
		g_r[i] = word;
	}
	g_f = 3;
	g_b = 0;
	for (int i = 0; i < 310; i++) (void)mgo_rand();
}

int mgo_rand(void)
{
	uint32_t v = (uint32_t)g_r[g_f] + (uint32_t)g_r[g_b];
	g_r[g_f] = (int32_t)v;
	if (++g_f >= 31) g_f = 0;
	if (++g_b >= 31) g_b = 0;
	return (int)(v >> 1);
}

static uint64_t splitmix64(uint64_t* s)
{
	uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

uint64_t mgo_chain_seed(uint64_t seed, uint64_t chain)
{
	uint64_t s = seed ^ (chain * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
	(void)splitmix64(&s);
	return s;
}

uint32_t mgo_chain_rand31(uint64_t* state) { return (uint32_t)(splitmix64(state) >> 33); }

typedef struct {
	int mode;
	uint64_t* state;
} Rng;

static uint32_t draw(Rng* r) { return r->mode == 0 ? (uint32_t)mgo_rand() : mgo_chain_rand31(r->state); }

/* ---- proposal generation (packet_slab_neighbour.c) ------------------------------------- */
typedef struct {
	size_t pos;
	MgoPacket old;
} Edit;

typedef struct {
	Edit* e;
	size_t count, cap;
} EditLog;

static void log_edit(EditLog* lg, size_t pos, MgoPacket old)
{
	if (lg->count == lg->cap) {
		lg->cap = lg->cap ? lg->cap * 2 : 64;
		lg->e = realloc(lg->e, sizeof(Edit) * lg->cap);
	}
	lg->e[lg->count].pos = pos;
	lg->e[lg->count].old = old;
	lg->count++;
}

/* packet_slab_neighbour.c:48-72 */
static int pick_from_topk(TopK* t, const Index* ix, const Model* m, MgoPacket* slab, Rng* rng, int best)
{
	find_candidates(t, ix, m, slab[m->pos]);
	size_t count = (size_t)t->count;
	if (count == 0) return 0;
	size_t choice = draw(rng) % count;
	for (int i = 1; i < 8; i++) {
		size_t c = draw(rng) % count;
		if (c > choice) choice = c;
	}
	if (draw(rng) % 8 == 0 || best) choice = count - 1;
	MgoPacket pk;
	while (topk_pop(t, &pk, NULL)) {
		slab[m->pos] = pk;
		if (choice-- == 0) return 1;
	}
	return 1;
}

/* packet_slab_neighbour.c:74-80 */
static int long_rep_ok(const Model* m, MgoPacket pk)
{
	return memcmp(m->data + m->pos - m->rep[pk.dist] - 1, m->data + m->pos, pk.len) == 0;
}

/* packet_slab_neighbour.c:119-152 */
static int mutate(const Model* m, TopK* t, const Index* ix, MgoPacket* slab, Rng* rng, EditLog* lg)
{
	size_t pos = m->pos;
	MgoPacket* first = &slab[pos];
	if (pos + 1 < m->n && draw(rng) % 2 == 0) {
		MgoPacket* second = &slab[pos + 1];
		if ((first->type == T_LONG_REP || first->type == T_MATCH) && first->len > 2) {
			log_edit(lg, pos, *first);
			log_edit(lg, pos + 1, *second);
			*second = *first;
			second->len--;
			*first = mk(T_LITERAL, 0, 1);
			return 1;
		} else if (first->type == T_LITERAL || first->type == T_SHORT_REP) {
			if (second->type == T_MATCH || second->type == T_LONG_REP) {
				size_t src = pos - second->dist;
				if (second->type == T_LONG_REP) src = pos - m->rep[second->dist];
				if (second->len < 273 && src > 0 && m->data[pos] == m->data[src - 1]) {
					log_edit(lg, pos, *first);
					*first = *second;
					first->len++;
					return 1;
				}
			}
		}
	}
	log_edit(lg, pos, slab[pos]);
	return pick_from_topk(t, ix, m, slab, rng, 0);
}

/* packet_slab_neighbour.c:82-117 */
static void repair(Model* m, Sink* s, TopK* t, const Index* ix, MgoPacket* slab, Rng* rng, EditLog* lg)
{
	size_t seen = 0;
	while (m->pos < m->n) {
		seen++;
		MgoPacket* pk = &slab[m->pos];
		MgoPacket before = *pk;
		if (pk->type == T_SHORT_REP || pk->type == T_LITERAL) {
			if (m->data[m->pos] == m->data[m->pos - m->rep[0] - 1]) {
				if (seen < 4) *pk = mk(T_SHORT_REP, 0, 1);
			} else {
				*pk = mk(T_LITERAL, 0, 1);
			}
		}
		if (pk->type == T_LONG_REP) {
			unsigned idx = 0;
			while (!long_rep_ok(m, *pk) && idx < 4) {
				pk->dist = idx;
				idx++;
			}
			if (!long_rep_ok(m, *pk)) pick_from_topk(t, ix, m, slab, rng, draw(rng) % 4 == 0);
		}
		if (!packet_eq(before, *pk)) log_edit(lg, m->pos, before);
		code_packet(m, s, *pk);
	}
}

size_t mgo_live_count(const MgoPacket* slab, size_t n)
{
	size_t pos = 0, count = 0;
	while (pos < n) {
		count++;
		pos += slab[pos].len;
	}
	return count;
}

/* packet_slab_neighbour.c:154-173 */
static int propose(const uint8_t* data, size_t n, MgoPacket* slab, TopK* t, const Index* ix, Rng* rng,
                   EditLog* lg, uint64_t* cost)
{
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_COST);
	size_t live = mgo_live_count(slab, n);
	size_t target = draw(rng) % live;
	for (size_t i = 0; i < target && m.pos < n; i++) code_packet(&m, &s, slab[m.pos]);
	if (!mutate(&m, t, ix, slab, rng, lg)) return 0;
	code_packet(&m, &s, slab[m.pos]);
	repair(&m, &s, t, ix, slab, rng, lg);
	*cost = s.cost;
	return 1;
}

long mgo_anneal_epoch(const uint8_t* data, size_t n, MgoPacket* slab, MgoPacket* best,
                      uint64_t* best_cost, int rng_mode, int reseed, unsigned seed,
                      uint64_t* rng_state, unsigned step, int num_iters, int first_eval, int evals,
                      long max_attempts, uint64_t* cur_cost_io, MgoTraceRec* trace, long trace_cap)
{
	Index ix;
	index_build(&ix, data, n);
	TopK t;
	topk_init(&t, 20);
	Rng rng = { rng_mode, rng_state };
	if (rng_mode == 0 && reseed) mgo_srand(seed);
	EditLog lg = { NULL, 0, 0 };
	uint64_t current = *cur_cost_io;
	long attempts = 0;
	int i = first_eval;
	while (i < first_eval + evals && attempts < max_attempts) {
		lg.count = 0;
		uint64_t cost = 0;
		int ok = propose(data, n, slab, &t, &ix, &rng, &lg, &cost);
		MgoTraceRec rec = { 0, 0, 0 };
		if (ok) {
			uint32_t r = draw(&rng);
			int uphill;
			if (rng_mode == 0) {
				/* main.c:86, evaluated with the reference's own C types */
				uphill = r % (i * i + 1 + step * num_iters / 2) < sqrt(num_iters);
			} else {
				uint64_t mod = (uint64_t)i * (uint64_t)i + 1 + (uint64_t)step * (uint64_t)num_iters / 2;
				uint64_t x = r % mod;
				uphill = x * x < (uint64_t)num_iters;
			}
			rec.cost = cost;
			rec.flags = 1;
			rec.undo_count = (uint32_t)lg.count;
			if (current == 0 || cost < current || uphill) {
				current = cost;
				rec.flags |= 2;
				if (*best_cost == 0 || current < *best_cost) {
					*best_cost = current;
					memcpy(best, slab, sizeof(MgoPacket) * n);
					rec.flags |= 4;
				}
			} else {
				while (lg.count > 0) { /* packet_slab_undo_stack.c:84-100: newest first */
					lg.count--;
					slab[lg.e[lg.count].pos] = lg.e[lg.count].old;
				}
			}
			i++;
		}
		if (trace != NULL && attempts < trace_cap) trace[attempts] = rec;
		attempts++;
	}
	*cur_cost_io = current;
	free(lg.e);
	topk_free(&t);
	index_free(&ix);
	return attempts;
}

/* ---- helpers for building test slabs --------------------------------------------------- */
void mgo_greedy_slab(const uint8_t* data, size_t n, MgoPacket* slab)
{
	Index ix;
	index_build(&ix, data, n);
	TopK t;
	topk_init(&t, 20);
	Model m;
	Sink s;
	model_init(&m, data, n);
	sink_init(&s, SINK_COST);
	for (size_t i = 0; i < n; i++) slab[i] = mk(T_LITERAL, 0, 1);
	while (m.pos < n) {
		MgoPacket none = mk(T_INVALID, 0, 0);
		find_candidates(&t, &ix, &m, none);
		MgoPacket pk = mk(T_LITERAL, 0, 1), cand;
		while (topk_pop(&t, &cand, NULL)) pk = cand;
		slab[m.pos] = pk;
		code_packet(&m, &s, pk);
	}
	topk_free(&t);
	index_free(&ix);
}

int mgo_slab_valid(const uint8_t* data, size_t n, const MgoPacket* slab)
{
	size_t pos = 0;
	uint32_t rep[4] = { 0, 0, 0, 0 };
	while (pos < n) {
		MgoPacket pk = slab[pos];
		if (pk.len == 0 || pos + pk.len > n) return 0;
		switch (pk.type) {
		case T_LITERAL:
			if (pk.len != 1) return 0;
			break;
		case T_SHORT_REP:
			if (pk.len != 1 || pos < (size_t)rep[0] + 1 || data[pos] != data[pos - rep[0] - 1]) return 0;
			break;
		case T_MATCH:
			if (pk.len < 2 || pk.len > 273 || pos < (size_t)pk.dist + 1) return 0;
			for (size_t i = 0; i < pk.len; i++)
				if (data[pos + i] != data[pos + i - pk.dist - 1]) return 0;
			rep[3] = rep[2];
			rep[2] = rep[1];
			rep[1] = rep[0];
			rep[0] = pk.dist;
			break;
		case T_LONG_REP: {
			if (pk.len < 2 || pk.len > 273 || pk.dist > 3) return 0;
			uint32_t d = rep[pk.dist];
			if (pos < (size_t)d + 1) return 0;
			for (size_t i = 0; i < pk.len; i++)
				if (data[pos + i] != data[pos + i - d - 1]) return 0;
			for (unsigned i = pk.dist; i > 0; i--) rep[i] = rep[i - 1];
			rep[0] = d;
			break;
		}
		default:
			return 0;
		}
		pos += pk.len;
	}
	return pos == n;
}
