// Device-side LZMA cost model for the annealing hot path (sm_100a).
//
// One warp owns one model.  A packet is priced in ONE step: every probability slot a packet
// touches belongs to a fixed lane ("slot class"), so the 9 modelled bits of a literal or the
// up-to-23 of a match are read, priced and adapted by different lanes at once.  This is exact
// because a packet never touches the same slot twice (reference
// src/lzma_packet_encoder.c:21-38,48-61,80-103,123-135) and a slot is only ever touched from
// its own lane, so the per-slot adaptation order is the reference's.
//
// The model lives in shared memory as a `Record` whose bytes are also the checkpoint format
// in HBM, so checkpoints move with single TMA bulk copies (cp.async.bulk, UBLKCP in SASS).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mg {

constexpr uint32_t FULL = 0xffffffffu;

// Packet types, reference src/lzma_packet.h:5-9
constexpr uint32_t T_INVALID = 0, T_LITERAL = 1, T_MATCH = 2, T_SHORT_REP = 3, T_LONG_REP = 4;

// ---- slot map at lc = lp = pb = 0 (only the reachable slots of src/lzma_state.h:15-55) ----
//
// Probabilities are stored PRE-SHIFTED (p << 2, the byte offset of p's entry in a transition-table section):
// a table address is then one three-input add, section + stored probability + table base.
//
// Plain literal tree (variant 0 of src/lzma_state.h:47-50) - the "ownership" block, 8 rows of 128 bytes:
// every slot a run of plain literals touches sits in a fixed shared-memory BANK, and in the literal fast
// path (walk_windows in mg_kernels.cuh) lane L is the only lane that touches bank L.  The lanes then share
// nothing: no bank conflicts on probabilities, no ordering between lanes, and all updates of one slot
// come from one lane in input order.
//   bank 0        is_match[0]                       (row 1; touched by every literal)
//   bank 1        tree depth 0 (the root)           (row 0, one slot per word)
//   banks 2-3     depth 1, banks 4-7 depth 2, banks 8-15 depth 3
//   banks 16-31   depth 4 (row 0), depth 5 (row 1), depth 6 (rows 2-3), depth 7 (rows 4-7); the bank is
//                 16 + four prefix bits XOR a per-depth constant, so that the nodes one byte value touches at
//                 depths 4..7 fall into different banks for the byte values that come in runs
//   row 0, odd slots: one spare slot per bank, always 0 - what a lane with nothing to do in a round
//                 "updates" (transition-table entry 0 maps probability 0 to itself at price 0)
//   rows 1-7, banks 0-15: the small per-state arrays (never touched by the fast path)
// Matched-literal trees (variants 1, 2) are in plain heap order.
constexpr uint32_t LEN_LOW = 2, LEN_MID = 10, LEN_HIGH = 18, LEN_SLOTS = 274;
constexpr uint32_t PROB_SHIFT = 2;
constexpr uint32_t PROB_INIT = 1024u << PROB_SHIFT;
constexpr uint32_t S_LIT0 = 0;                      // 512 slot spaces, see lit0_slot()
constexpr uint32_t S_ISMATCH = 64;                  // [12]; is_match[0] in bank 0
constexpr uint32_t S_ISREP = 128;
constexpr uint32_t S_ISREPG0 = S_ISREP + 12;
constexpr uint32_t S_ISREPG1 = 192;
constexpr uint32_t S_ISREPG2 = S_ISREPG1 + 12;
constexpr uint32_t S_ISREP0LONG = 256;
constexpr uint32_t S_LITV = 512;                    // [2 variants][256] heap order, node 0 unused
constexpr uint32_t S_LEN = S_LITV + 512;            // choice1, choice2, low[8], mid[8], high[256]
constexpr uint32_t S_REPLEN = S_LEN + LEN_SLOTS;
constexpr uint32_t S_POSSLOT = S_REPLEN + LEN_SLOTS;  // [4][64]
constexpr uint32_t S_ALIGN = S_POSSLOT + 256;       // [16]
constexpr uint32_t S_POSCODER = S_ALIGN + 16;       // [115]
constexpr uint32_t S_TOTAL = S_POSCODER + 115;
constexpr uint32_t S_COUNT = ((S_TOTAL + 7) / 8) * 8;  // probabilities stored per record (16-byte multiple)
constexpr uint32_t S_DUMMY = 1;                     // bank 0's spare slot: the address lanes without a slot class point at (never coded)
// per-depth XOR constants of the deep levels (bits of the prefix that select the bank), chosen by simulating
// the lanes' queue lengths over the 1 MiB mixed corpus: runs of 0x00, 0xff, ' ', '0', '-', '=', 0x90, 0xcc ...
// keep their four deep nodes in four different banks
constexpr uint32_t LITX4 = 0, LITX5 = 13u << 1, LITX6 = 2u << 1, LITX7 = 11u << 1;
static_assert(S_ISMATCH % 64 == 0 && S_ISREP % 64 == 0 && S_ISREPG1 % 64 == 0 && S_ISREP0LONG % 64 == 0, "small arrays sit in banks 0-15");

// Slot of node (1 << depth | prefix) of the plain literal tree.
__host__ __device__ inline uint32_t lit0_slot(uint32_t depth, uint32_t prefix)
{
	if (depth <= 3) return 2u * ((1u << depth) | prefix);
	if (depth == 4) return 2u * (16u + (prefix ^ LITX4));
	const uint32_t u = prefix ^ (depth == 5 ? LITX5 : depth == 6 ? LITX6 : LITX7);
	if (depth == 5) return 64u + 32u + u;
	if (depth == 6) return 128u + 32u + u + (u & 32u);
	return 256u + 32u + u + (u & 0x60u);
}

// Slot of literal-tree node `node` = variant << 8 | (1 << depth | prefix), the reference's index
// into its 0x300 literal probabilities (src/lzma_packet_encoder.c:106-136).
__host__ __device__ inline uint32_t lit_slot(uint32_t node)
{
	const uint32_t v = node >> 8, i = node & 0xff;
	if (v != 0) return S_LITV + (v - 1) * 256 + i;
	if (i == 0) return S_DUMMY;
	uint32_t d = 0;
	while ((i >> (d + 1)) != 0) d++;
	return S_LIT0 + lit0_slot(d, i - (1u << d));
}

// Working set of one warp == checkpoint record in HBM.
struct alignas(16) Record {
	uint16_t probs[S_COUNT];
	uint32_t rep[4];
	uint32_t pos;   // byte position of the next packet
	uint32_t pidx;  // index of that packet in the live chain
	uint32_t ctx;   // the 12-state automaton
	uint32_t pad0;
	uint64_t cost;  // cost of everything before pos, 1/2048 bit
	uint64_t pad1;
};
static_assert(sizeof(Record) % 16 == 0 && sizeof(Record) == S_COUNT * 2 + 48, "record must stay a multiple of 16 bytes for cp.async.bulk");

// Packed slab slot: dist[31:0] | len[47:32] | type[55:48]
__host__ __device__ __forceinline__ uint64_t pk_pack(uint32_t type, uint32_t dist, uint32_t len)
{
	return (uint64_t)dist | ((uint64_t)len << 32) | ((uint64_t)type << 48);
}
__host__ __device__ __forceinline__ uint32_t pk_type(uint64_t p) { return (uint32_t)(p >> 48) & 0xff; }
__host__ __device__ __forceinline__ uint32_t pk_len(uint64_t p) { return (uint32_t)(p >> 32) & 0xffff; }
__host__ __device__ __forceinline__ uint32_t pk_dist(uint64_t p) { return (uint32_t)p; }
constexpr uint64_t PK_LITERAL = (1ull << 48) | (1ull << 32);
constexpr uint64_t PK_SHORT_REP = (3ull << 48) | (1ull << 32);

// Registers every lane of the warp holds identically.
struct Model {
	uint32_t pos, pidx, ctx;
	uint32_t rep0, rep1, rep2, rep3;
};

// ---- small PTX helpers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t done;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.b32 %0, 1, 0, p;\n\t}"
	    : "=r"(done)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return done != 0;
}
// global -> shared bulk copy (TMA), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
// shared -> global bulk copy (TMA), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
	             "r"(smem_u32(smem_src)), "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the 12-state automaton, src/lzma_state.c:29-57 ----------------------------------------
__device__ __forceinline__ uint32_t next_ctx(uint32_t ctx, uint32_t type)
{
	if (type == T_LITERAL) return ctx < 4 ? 0 : (ctx < 10 ? ctx - 3 : ctx - 6);
	if (type == T_MATCH) return ctx < 7 ? 7 : 10;
	if (type == T_SHORT_REP) return ctx < 7 ? 9 : 11;
	return ctx < 7 ? 8 : 11;
}

// Shared-memory accessors with 32-bit shared addresses (LDS/STS without generic-address math).
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
	uint16_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
	asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}

// Arrays in shared memory addressed by their 32-bit shared address.  Going through generic
// pointers kept in structs made the compiler re-derive the shared window for every access
// (8 % of the instructions executed on match-heavy slabs were cvta sequences).
struct SmemU16 {
	uint32_t a;
	__device__ __forceinline__ uint32_t get(uint32_t i) const { return lds_u16(a + 2 * i); }
	__device__ __forceinline__ void set(uint32_t i, uint32_t v) const { sts_u16(a + 2 * i, v); }
};
struct SmemU32 {
	uint32_t a;
	__device__ __forceinline__ uint32_t get(uint32_t i) const { return lds_u32(a + 4 * i); }
	__device__ __forceinline__ void set(uint32_t i, uint32_t v) const { sts_u32(a + 4 * i, v); }
};

// Transition tables, six sections of 2048 x u32 (8 KB each), indexed by the stored probability (p << 2 = the
// entry's byte offset inside a section):
//   section 0, 1      one step, bit 0 / 1
//   section 2 + ab    two steps on one slot, bits (a, b) in that order
//   low 16 bits  = the probability after the step(s), stored form   (src/probability_model.c:5-15)
//   high 16 bits = the price of the step(s) (src/perplexity_encoder.c:6-10; two steps: at most 45 056)
// One shared-memory load replaces the price lookup plus the shift/add/select update.
constexpr uint32_t TRANS_SECTION_SHIFT = 13, TRANS_SECTIONS = 6, TRANS_WORDS = TRANS_SECTIONS * 2048;
__device__ __forceinline__ void code_bit(SmemU16 probs, SmemU32 trans, uint32_t slot, uint32_t bit, uint32_t& acc)
{
	const uint32_t t = lds_u32(trans.a + (bit << TRANS_SECTION_SHIFT) + probs.get(slot));
	acc += t >> 16;
	probs.set(slot, t);
}

__device__ __forceinline__ uint32_t bit_price(SmemU16 probs, SmemU32 trans, uint32_t slot, uint32_t bit)
{
	return lds_u32(trans.a + (bit << TRANS_SECTION_SHIFT) + probs.get(slot)) >> 16;
}

// ---- lane -> (slot, bit) maps ----------------------------------------------------------------
// Every lane has ONE role per packet kind, fixed at kernel start, and the maps below are pure
// select arithmetic (no divergent branches): all lanes run the same instructions and the one
// probability update that follows is issued once for the whole warp.  Lane order == the
// reference's coding order, which the range-coder kernel relies on.
//   LITERAL   : 0 is_match | 1..8 literal tree depth 0..7
//   otherwise : 0 is_match | 1 is_rep | 2 is_rep_g0 | 3 is_rep0_long or is_rep_g1 | 4 is_rep_g2
//               | 5 len choice_1 | 6 len choice_2 | 7..14 length tree | 15..20 pos-slot tree
//               | 21..25 reverse tree (pos_coder or align; direct bits sit just before lane 21)
// A slot is only ever touched from one lane: is_match 0, is_rep 1, ... length trees 5..14 (the match
// and rep length coders are different slots), distance 15..25, literal tree 1..8.
enum LaneGroup { G_HEADER = 0, G_CHOICE = 1, G_LENTREE = 2, G_SLOTTREE = 3, G_REVTREE = 4, G_NONE = 5 };
constexpr int FIRST_REVTREE_LANE = 21;

struct LaneRole {
	uint32_t grp;  // LaneGroup for non-literal packets
	uint32_t t;    // index inside the group
};

__device__ __forceinline__ LaneRole lane_role(int lane)
{
	LaneRole r;
	if (lane < 5) {
		r.grp = G_HEADER;
		r.t = (uint32_t)lane;
	} else if (lane < 7) {
		r.grp = G_CHOICE;
		r.t = (uint32_t)lane - 5;
	} else if (lane < 15) {
		r.grp = G_LENTREE;
		r.t = (uint32_t)lane - 7;
	} else if (lane < 21) {
		r.grp = G_SLOTTREE;
		r.t = (uint32_t)lane - 15;
	} else if (lane < 26) {
		r.grp = G_REVTREE;
		r.t = (uint32_t)lane - 21;
	} else {
		r.grp = G_NONE;
		r.t = 0;
	}
	return r;
}

struct DistParts {
	uint32_t pslot, nlow, low, rbase, rbits;  // rbits: bits coded by the reverse tree (0 when dist < 4)
	uint32_t direct;                          // number of direct bits
};

// src/lzma_packet_encoder.c:71-104
__device__ __forceinline__ DistParts dist_parts(uint32_t dist)
{
	DistParts d;
	if (dist < 4) {
		d.pslot = dist;
		d.nlow = d.low = d.rbase = d.rbits = d.direct = 0;
		return d;
	}
	d.nlow = 30u - (uint32_t)__clz(dist);
	d.low = dist & ((1u << d.nlow) - 1);
	uint32_t high = dist >> d.nlow;
	d.pslot = d.nlow * 2 + high;
	if (d.pslot < 14) {
		d.rbase = S_POSCODER + (high << d.nlow) - d.pslot;
		d.rbits = d.nlow;
		d.direct = 0;
	} else {
		d.rbase = S_ALIGN;
		d.rbits = 4;
		d.direct = d.nlow - 4;
		d.low &= 15;  // the part the align tree codes; the direct bits carry no cost state
	}
	return d;
}

// Literal tree (src/lzma_packet_encoder.c:106-136): depth 0..7.
__device__ __forceinline__ void lit_event(uint32_t depth, uint32_t byte, bool matched_mode, uint32_t mbyte,
                                          uint32_t& slot, uint32_t& bit)
{
	const uint32_t top = byte >> (8 - depth);
	bit = (byte >> (7 - depth)) & 1;
	uint32_t v = 0;
	if (matched_mode && (mbyte >> (8 - depth)) == top) v = 1u + ((mbyte >> (7 - depth)) & 1);
	slot = v == 0 ? S_LIT0 + lit0_slot(depth, top) : S_LITV + (v - 1) * 256 + ((1u << depth) | top);
}

// The length coder's three trees (src/lzma_packet_encoder.c:42-63) as one: nb bits of value w
// under `tree`, after choice_1 / choice_2.
struct LenParts {
	uint32_t v, nb, w, tree;
};
__device__ __forceinline__ LenParts len_parts(uint32_t base, uint32_t len)
{
	LenParts l;
	l.v = len - 2;
	l.nb = l.v < 16 ? 3u : 8u;
	l.w = l.v < 8 ? l.v : (l.v < 16 ? l.v - 8 : l.v - 16);
	l.tree = base + (l.v < 8 ? LEN_LOW : l.v < 16 ? LEN_MID : LEN_HIGH);
	return l;
}

// MATCH (src/lzma_packet_encoder.c:138-146): header, length, distance.
__device__ __forceinline__ bool match_event(const LaneRole r, uint32_t ctx, uint32_t len, const DistParts& d,
                                            uint32_t& slot, uint32_t& bit)
{
	const LenParts l = len_parts(S_LEN, len);
	const uint32_t t = r.t;
	// header: is_match = 1, is_rep = 0 (lanes 2..4 idle)
	const uint32_t s_h = (t == 0 ? S_ISMATCH : S_ISREP) + ctx;
	const uint32_t b_h = t == 0;
	const bool a_h = t < 2;
	// choice_1 / choice_2
	const uint32_t s_c = S_LEN + t;
	const uint32_t b_c = t == 0 ? l.v >= 8 : l.v >= 16;
	const bool a_c = t == 0 || l.v >= 8;
	// length tree
	const uint32_t sh = l.nb > t ? l.nb - t : 0;
	const uint32_t s_l = l.tree + ((1u << t) | (l.w >> sh));
	const uint32_t b_l = (l.w >> (sh ? sh - 1 : 0)) & 1;
	const bool a_l = t < l.nb;
	// pos-slot tree under the length context
	const uint32_t lctx = l.v < 3 ? l.v : 3;
	const uint32_t ts = t < 5 ? t : 5;
	const uint32_t s_s = S_POSSLOT + lctx * 64 + ((1u << ts) | (d.pslot >> (6 - ts)));
	const uint32_t b_s = (d.pslot >> (5 - ts)) & 1;
	// reverse tree: pos_coder for short distances, align for long ones
	const uint32_t prefix = t ? (__brev(d.low & ((1u << t) - 1)) >> (32 - t)) : 0;
	const uint32_t s_r = d.rbase + ((1u << t) | prefix);
	const uint32_t b_r = (d.low >> t) & 1;
	const bool a_r = t < d.rbits;
	const uint32_t g = r.grp;
	slot = g == G_HEADER ? s_h : g == G_CHOICE ? s_c : g == G_LENTREE ? s_l : g == G_SLOTTREE ? s_s : s_r;
	bit = g == G_HEADER ? b_h : g == G_CHOICE ? b_c : g == G_LENTREE ? b_l : g == G_SLOTTREE ? b_s : b_r;
	return g == G_HEADER ? a_h : g == G_CHOICE ? a_c : g == G_LENTREE ? a_l : g == G_SLOTTREE ? true : g == G_REVTREE ? a_r : false;
}

// SHORT_REP / LONG_REP (src/lzma_packet_encoder.c:148-167): header bits, then the rep length coder.
__device__ __forceinline__ bool rep_event(const LaneRole r, uint32_t type, uint32_t ctx, uint32_t len, uint32_t idx,
                                          uint32_t& slot, uint32_t& bit)
{
	const bool is_long = type == T_LONG_REP;
	const LenParts l = len_parts(S_REPLEN, is_long ? len : 2);
	const uint32_t t = r.t;
	// header lanes: is_match 1 | is_rep 1 | g0 | rep0_long (short rep, long rep 0) or g1 | g2
	const bool rep0 = !is_long || idx == 0;
	const uint32_t base_h = t == 0 ? S_ISMATCH : t == 1 ? S_ISREP : t == 2 ? S_ISREPG0 : t == 3 ? (rep0 ? S_ISREP0LONG : S_ISREPG1) : S_ISREPG2;
	const uint32_t s_h = base_h + ctx;
	const uint32_t b_h = t < 2 ? 1u : t == 2 ? (is_long && idx != 0) : t == 3 ? (is_long && idx != 1) : idx != 2;
	const bool a_h = t < 4 ? true : (is_long && idx >= 2);
	const uint32_t s_c = S_REPLEN + t;
	const uint32_t b_c = t == 0 ? l.v >= 8 : l.v >= 16;
	const bool a_c = is_long && (t == 0 || l.v >= 8);
	const uint32_t sh = l.nb > t ? l.nb - t : 0;
	const uint32_t s_l = l.tree + ((1u << t) | (l.w >> sh));
	const uint32_t b_l = (l.w >> (sh ? sh - 1 : 0)) & 1;
	const bool a_l = is_long && t < l.nb;
	const uint32_t g = r.grp;
	slot = g == G_HEADER ? s_h : g == G_CHOICE ? s_c : s_l;
	bit = g == G_HEADER ? b_h : g == G_CHOICE ? b_c : b_l;
	return g == G_HEADER ? a_h : g == G_CHOICE ? a_c : g == G_LENTREE ? a_l : false;
}

// Lane's event for a whole packet of any type (the walk's general path, the range coder).
__device__ __forceinline__ bool packet_event(int lane, const LaneRole r, uint32_t type, uint32_t len, uint32_t dist,
                                             uint32_t ctx, uint32_t byte, uint32_t mbyte, const DistParts& dp,
                                             uint32_t& slot, uint32_t& bit)
{
	if (type == T_MATCH) return match_event(r, ctx, len, dp, slot, bit);  // uniform branches: the type is
	if (type != T_LITERAL) return rep_event(r, type, ctx, len, dist, slot, bit);  // the same on every lane
	const uint32_t depth = lane >= 1 && lane <= 8 ? (uint32_t)lane - 1 : 0;
	lit_event(depth, byte, ctx >= 7, mbyte, slot, bit);
	if (lane == 0) {
		slot = S_ISMATCH + ctx;
		bit = 0;
	}
	return lane <= 8;
}

// ---- MATCH fast path ---------------------------------------------------------------------------
// The events of a MATCH (src/lzma_packet_encoder.c:138-146) from two packed words that are the same
// on every lane and four per-lane constants (a 512-byte table in shared memory):
//   F  tree fields, each left-aligned in its byte (most significant bit first, the order the trees
//      consume them): byte3 length tree, byte2 pos slot, byte1 reverse tree (bit-reversed low bits),
//      byte0 bit0 is_match=1, bit1 is_rep=0, bit2 choice_1, bit3 choice_2
//   G  the variable part of each group's slot base: byte0 ctx, byte1 length tree offset,
//      byte2 64 * length context, byte3 reverse-tree base - S_ALIGN
// A lane at depth t of its tree reads the top t+1 bits of its field: prefix and bit in one shift.
struct LaneConst {
	uint32_t shb;    // F >> shb puts the lane's bit at bit 0 and its prefix above it
	uint32_t mask2;  // (2 << t) - 1
	uint32_t base;   // fixed part of the slot index, tree root offset (1 << t) included
	uint32_t sel;    // __byte_perm selector of the lane's byte of G (0x4444 = none)
};

__host__ __device__ inline LaneConst match_lane_const(int lane)
{
	LaneConst c = {0, 0, S_DUMMY, 0x4444};
	if (lane == 0) c = {0, 1, S_ISMATCH, 0x4440};
	else if (lane == 1) c = {1, 1, S_ISREP, 0x4440};
	else if (lane == 5) c = {2, 1, S_LEN, 0x4444};
	else if (lane == 6) c = {3, 1, S_LEN + 1, 0x4444};
	else if (lane >= 7 && lane < 15) {
		const uint32_t t = (uint32_t)lane - 7;
		c = {24 + 7 - t, (2u << t) - 1, S_LEN + (1u << t), 0x4441};
	} else if (lane >= 15 && lane < 21) {
		const uint32_t t = (uint32_t)lane - 15;
		c = {16 + 7 - t, (2u << t) - 1, S_POSSLOT + (1u << t), 0x4442};
	} else if (lane >= 21 && lane < 26) {
		const uint32_t t = (uint32_t)lane - 21;
		c = {8 + 7 - t, (2u << t) - 1, S_ALIGN + (1u << t), 0x4443};
	}
	return c;
}

struct MatchDesc {
	uint32_t F, G, amask, direct;
};

__device__ __forceinline__ MatchDesc match_desc(uint32_t ctx, uint32_t len, uint32_t dist)
{
	MatchDesc d;
	const uint32_t v = len - 2;
	const uint32_t cls = (v >= 8 ? 1u : 0u) + (v >= 16 ? 1u : 0u);      // low / mid / high tree
	const uint32_t w = v - (cls == 0 ? 0u : cls == 1 ? 8u : 16u);
	const uint32_t fieldL = cls == 2 ? w : w << 5;
	const uint32_t lenoff = LEN_LOW + 8 * cls;
	const uint32_t lctx = v < 3 ? v : 3;
	// src/lzma_packet_encoder.c:71-104
	uint32_t pslot = dist, low = 0, rbase = S_ALIGN, rbits = 0, direct = 0;
	if (dist >= 4) {
		const uint32_t nlow = 30u - (uint32_t)__clz(dist);
		const uint32_t high = dist >> nlow;
		low = dist & ((1u << nlow) - 1);
		pslot = nlow * 2 + high;
		if (pslot < 14) {
			rbase = S_POSCODER + (high << nlow) - pslot;
			rbits = nlow;
		} else {
			rbits = 4;
			direct = nlow - 4;
			low &= 15;
		}
	}
	const uint32_t fieldR = __brev(low) >> 24;
	d.F = (fieldL << 24) | (pslot << 18) | (fieldR << 8) | 1u | (cls >= 1 ? 4u : 0u) | (cls == 2 ? 8u : 0u);
	d.G = ctx | (lenoff << 8) | (lctx << 22) | ((rbase - S_ALIGN) << 24);
	d.amask = 0x23u | (cls >= 1 ? 0x40u : 0u) | ((cls == 2 ? 0xffu : 0x7u) << 7) | (0x3fu << 15) | (((1u << rbits) - 1u) << 21);
	d.direct = direct;
	return d;
}

// src/lzma_state.c:59-81 + src/lzma_packet_encoder.c:192-193
__device__ __forceinline__ void model_advance(Model& m, uint32_t type, uint32_t len, uint32_t dist)
{
	if (type == T_MATCH) {
		m.rep3 = m.rep2;
		m.rep2 = m.rep1;
		m.rep1 = m.rep0;
		m.rep0 = dist;
	} else if (type == T_LONG_REP) {
		uint32_t d = dist == 0 ? m.rep0 : dist == 1 ? m.rep1 : dist == 2 ? m.rep2 : m.rep3;
		if (dist > 2) m.rep3 = m.rep2;
		if (dist > 1) m.rep2 = m.rep1;
		if (dist > 0) m.rep1 = m.rep0;
		m.rep0 = d;
	}
	m.ctx = next_ctx(m.ctx, type);
	m.pos += len;
	m.pidx += 1;
}

__device__ __forceinline__ uint32_t model_rep(const Model& m, uint32_t idx)
{
	return idx == 0 ? m.rep0 : idx == 1 ? m.rep1 : idx == 2 ? m.rep2 : m.rep3;
}

// Price + adapt one packet across the warp.  `byte` = data[pos]; mbyte only read when the packet
// is a literal in matched mode.  Returns the number of modelled bits (uniform).
__device__ __forceinline__ uint32_t apply_packet(int lane, SmemU16 probs, SmemU32 trans, Model& m,
                                                 uint32_t type, uint32_t len, uint32_t dist, uint32_t byte,
                                                 uint32_t mbyte, uint32_t& acc)
{
	DistParts dp;
	dp.pslot = dp.nlow = dp.low = dp.rbase = dp.rbits = dp.direct = 0;
	if (type == T_MATCH) dp = dist_parts(dist);
	uint32_t slot = 0, bit = 0;
	bool active = packet_event(lane, lane_role(lane), type, len, dist, m.ctx, byte, mbyte, dp, slot, bit);
	// Re-converge before the shared tail: without this the compiler clones the probability
	// update and the model bookkeeping into every divergent lane group (measured: that tail ran
	// with ~3 active lanes, about ten times per packet).
	__syncwarp();
	if (active) code_bit(probs, trans, slot, bit, acc);
	if (lane == 0) acc += dp.direct << 11;  // src/perplexity_encoder.c:12-17
	model_advance(m, type, len, dist);
	return __popc(__ballot_sync(FULL, active));
}

// `half` = the stored form of probability 1/2 (PROB_INIT in the scoring kernels, 1024 in the range coder);
// the spare slots of row 0 stay 0
__device__ __forceinline__ void model_init(int lane, SmemU16 probs, Model& m, uint32_t half = PROB_INIT)
{
	for (uint32_t i = (uint32_t)lane; i < S_COUNT; i += 32) probs.set(i, (i < 64 && (i & 1)) ? 0u : half);
	m.pos = m.pidx = m.ctx = 0;
	m.rep0 = m.rep1 = m.rep2 = m.rep3 = 0;
	__syncwarp();
}

// Window of 32 consecutive slab slots + data bytes held across the warp's registers (one slot per
// lane), so the live-chain walk costs one shuffle per packet instead of a dependent load; the next
// window is staged in shared memory by cp.async (window_prefetch).
// meta = type[2:0] | len[15:3] | data byte[23:16]; dist separately (only matches and long reps need it).
constexpr uint32_t META_LITERAL = T_LITERAL | (1u << 3);
__device__ __forceinline__ uint32_t meta_type(uint32_t meta) { return meta & 7; }
__device__ __forceinline__ uint32_t meta_len(uint32_t meta) { return (meta >> 3) & 0x1fff; }
__device__ __forceinline__ uint32_t meta_byte(uint32_t meta) { return meta >> 16; }

struct Window {
	uint32_t meta, dist;  // current window, decoded
	uint32_t base;        // multiple of 32; WINDOW_NONE = nothing loaded
	uint32_t pf_base;     // window being copied into the warp's staging area by cp.async (WINDOW_NONE = none)
	uint32_t litmask;     // bit i: slot base+i is a canonical LITERAL
	uint32_t md_base;     // window whose MATCH descriptors are mirrored in shared memory (see window_matches)
};
constexpr uint32_t WINDOW_NONE = 0x7fffffe0u;  // never within 32 of a real position (inputs < 2 GiB... see mg_ctx_create)

__device__ __forceinline__ void window_decode(uint32_t lo, uint32_t hi, uint32_t byte, uint32_t& meta, uint32_t& dist)
{
	const uint32_t len = hi & 0xffffu, type = (hi >> 16) & 0xffu;
	meta = type | ((len < 0x1fff ? len : 0x1fff) << 3) | (byte << 16);
	dist = lo;
}

// The next window travels global -> shared memory by cp.async (LDGSTS): no register holds it and
// nothing waits for it until the walk gets there, a whole window of work later.  (Held in registers,
// the raw loads were spilled by the compiler - i.e. consumed at once - and every window stalled on
// HBM latency: 11 % of all warp samples.)  `stage` = shared address of 32 x 8 B slots + 32 data bytes.
__device__ __forceinline__ void window_prefetch(int lane, const uint64_t* __restrict__ slab,
                                                const uint8_t* __restrict__ data, uint32_t n, uint32_t base, uint32_t stage)
{
	const uint32_t i = base + (uint32_t)lane;
	const uint32_t in = i < n ? 8u : 0u;  // src-size 0 zero-fills
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(stage + 8u * (uint32_t)lane), "l"(slab + (i < n ? i : 0)), "r"(in) : "memory");
	if (lane < 8) {
		// the input buffer is padded by 32 bytes (mg_ctx_create), base is a multiple of 32
		const uint32_t b = base + 4u * (uint32_t)lane;
		const uint32_t inb = b < n ? 4u : 0u;
		asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(stage + 256u + 4u * (uint32_t)lane), "l"(data + (b < n ? b : 0)), "r"(inb) : "memory");
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void window_seek(int lane, Window& w, const uint64_t* __restrict__ slab,
                                            const uint8_t* __restrict__ data, uint32_t n, uint32_t pos, uint32_t stage)
{
	const uint32_t want = pos & ~31u;
	if (want == w.base) return;
	uint32_t lo, hi, byte;
	// also on a jump: a copy still in flight must not land on top of the one issued below
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	if (want == w.pf_base) {
		__syncwarp();
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(stage + 8u * (uint32_t)lane) : "memory");
		byte = lds_u8(stage + 256u + (uint32_t)lane);
		if (want + (uint32_t)lane >= n) byte = 0;  // the padding behind the input is not part of it
	} else {
		// a jump: nothing staged for this window
		const uint32_t i = want + (uint32_t)lane;
		const bool in = i < n;
		const uint64_t pk = in ? slab[i] : 0;
		byte = in ? data[i] : 0;
		lo = (uint32_t)pk;
		hi = (uint32_t)(pk >> 32);
	}
	window_decode(lo, hi, byte, w.meta, w.dist);
	w.base = want;
	__syncwarp();  // every lane has read the staging area
	window_prefetch(lane, slab, data, n, want + 32, stage);
	w.pf_base = want + 32;
	w.litmask = __ballot_sync(FULL, (w.meta & 0xffffu) == META_LITERAL);
}

__device__ __forceinline__ uint32_t window_meta(const Window& w, uint32_t pos)
{
	return __shfl_sync(FULL, w.meta, (int)(pos - w.base));
}
__device__ __forceinline__ uint32_t window_dist(const Window& w, uint32_t pos)
{
	return __shfl_sync(FULL, w.dist, (int)(pos - w.base));
}

// Checkpoint traffic: one TMA bulk copy each way, issued by lane 0.
__device__ __forceinline__ void record_store(int lane, Record* rec, const Model& m, uint64_t cost, Record* dst)
{
	__syncwarp();
	if (lane == 0) {
		rec->ctx = m.ctx;
		rec->rep[0] = m.rep0;
		rec->rep[1] = m.rep1;
		rec->rep[2] = m.rep2;
		rec->rep[3] = m.rep3;
		rec->pos = m.pos;
		rec->pidx = m.pidx;
		rec->cost = cost;
		fence_async_smem();
		bulk_s2g(dst, rec, (uint32_t)sizeof(Record));
		bulk_commit();
		bulk_wait_read();
	}
	__syncwarp();
}

__device__ __forceinline__ void record_load(int lane, Record* rec, Model& m, uint64_t& cost, const Record* src,
                                            uint64_t* bar, uint32_t& parity)
{
	__syncwarp();
	if (lane == 0) {
		mbar_expect_tx(bar, (uint32_t)sizeof(Record));
		bulk_g2s(rec, src, (uint32_t)sizeof(Record), bar);
	}
	while (!mbar_try_wait(bar, parity)) {
	}
	parity ^= 1;
	__syncwarp();
	m.ctx = rec->ctx;
	m.rep0 = rec->rep[0];
	m.rep1 = rec->rep[1];
	m.rep2 = rec->rep[2];
	m.rep3 = rec->rep[3];
	m.pos = rec->pos;
	m.pidx = rec->pidx;
	cost = rec->cost;
	__syncwarp();
}

// splitmix64, the chains' counter-based generator (top 31 bits, like rand()).
__device__ __forceinline__ uint32_t rng31(uint64_t& s)
{
	s += 0x9E3779B97F4A7C15ull;
	uint64_t z = s;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (uint32_t)(z >> 33);
}

}  // namespace mg
