// Warp-level match finder + exact top-k (sm_100a).
//
// Replaces substring_enumerator_for_each -> packet_enumerator -> top_k_packet_finder ->
// max_heap of the reference.  32 earlier occurrences of the bigram at `pos` are examined per
// step, one per lane: each lane extends its match, prices all of its candidates from the
// pre-state (a packet's price is a pure function of the model before it, so no state copy
// per candidate), and only lanes that could still enter the heap replay their candidates,
// in the reference's enumeration order, into a bounded max-heap kept in shared memory with
// the reference's exact tie behaviour (src/max_heap.c:82-121, `<=` eviction
// src/top_k_packet_finder.c:89).  Once the heap is full its root never increases, so
// skipping a lane whose cheapest candidate is above the root is exact.
#pragma once
#include "mg_device.cuh"

namespace mg {

constexpr int MAX_K = 32;
// Optional limits beyond the reference (mg_ctx_set_finder_limits): 0 = none.
struct FindLimits {
	uint32_t window;   // farthest match start, bytes before the position
	uint32_t max_occ;  // only the nearest max_occ earlier occurrences of the bigram
};
constexpr uint32_t FIND_GAVE_UP = 0xffffffffu;  // warp_find: the deadline passed inside a long bucket
constexpr uint32_t MAX_MATCH = 273;  // src/packet_enumerator.c:6-7
#ifndef MG_FINDER_PM
#define MG_FINDER_PM 1  // 0 switches the running-maximum filter off (kernel experiments)
#endif

struct FindScratch {
	uint64_t ent_pk[MAX_K];
	uint32_t ent_price[MAX_K];
	// [0] match lengths, [1] rep lengths; index len-2.  Only live during a find: between finds the
	// walk keeps its window staging here (see below).
	uint32_t len_price[2][MAX_MATCH - 1];
	uint32_t hkey[MAX_K];                  // heap order -> price << 5 | entry index
	uint8_t pop_order[MAX_K];              // entry indices, worst first
	uint32_t count;
	uint32_t pops;
	uint32_t candidates;  // candidates enumerated by the last find
	uint32_t chunks;      // 32-occurrence steps the last find took (its cost in the step budget)
};

// Between finds the length-price tables double as the walk's staging area (mg_kernels.cuh):
//   [0, 512)     one 16-byte MATCH descriptor per slot of the window (window_matches())
//   [512, 800)   the NEXT window's 32 slab slots (8 B) + 32 data bytes, filled by cp.async (window_prefetch())
constexpr uint32_t QUEUE_ROUNDS = 16;
constexpr uint32_t MATCH_DESC_OFFSET = 0;
constexpr uint32_t STAGE_OFFSET = MATCH_DESC_OFFSET + 32 * 16;
constexpr uint32_t STAGE_BYTES = 32 * 8 + 32;
static_assert(STAGE_OFFSET % 16 == 0 && STAGE_OFFSET + STAGE_BYTES <= sizeof(uint32_t) * 2 * (MAX_MATCH - 1), "window mirrors must fit in len_price");

// floor(c / len) for c*len < 2^32 via one multiply: recip[len] = floor((2^32-1)/len) + 1
__device__ __forceinline__ uint32_t per_byte(uint32_t cost, uint32_t len, SmemU32 recip)
{
	return len == 1 ? cost : __umulhi(cost, recip.get(len));
}

// The heap order lives in hkey[]: heap position -> (price << 5) | entry index, so a sift compares
// one word per node (prices are below 2^20: a packet costs at most ~5.5e5).
// "Hole" form of the reference's swap loops: the moving element is written once, at its final place.

// src/max_heap.c:82-105 (left child preferred on ties, strict comparisons)
__device__ __forceinline__ void heap_sink(FindScratch* fs, uint32_t p, uint32_t key, uint32_t count)
{
	const uint32_t price = key >> 5;
	for (;;) {
		const uint32_t l = 2 * p + 1;
		if (l >= count) break;
		const uint32_t kl = fs->hkey[l];
		const uint32_t kr = l + 1 < count ? fs->hkey[l + 1] : 0u;
		uint32_t big = l, kb = kl;
		if ((kr >> 5) > (kl >> 5)) {
			big = l + 1;
			kb = kr;
		}
		if ((kb >> 5) <= price) break;
		fs->hkey[p] = kb;
		p = big;
	}
	fs->hkey[p] = key;
}

// src/max_heap.c:107-121
__device__ __forceinline__ void heap_rise(FindScratch* fs, uint32_t p, uint32_t key)
{
	const uint32_t price = key >> 5;
	while (p > 0) {
		const uint32_t parent = (p - 1) / 2;
		const uint32_t kp = fs->hkey[parent];
		if (price <= (kp >> 5)) break;
		fs->hkey[p] = kp;
		p = parent;
	}
	fs->hkey[p] = key;
}

// src/top_k_packet_finder.c:72-93,99-101.  Called by exactly one lane at a time.  Deliberately
// out of line: it is reached from a dozen places in the finder, and one copy keeps the finder
// (and the annealing kernel around it) inside the instruction cache.
__device__ __noinline__ void heap_offer(FindScratch* fs, uint32_t k, uint64_t pk, uint32_t price, uint64_t excluded)
{
	if (pk == excluded) return;
	const uint32_t count = fs->count;
	if (count < k) {
		fs->ent_pk[count] = pk;
		fs->ent_price[count] = price;
		fs->count = count + 1;
		heap_rise(fs, count, (price << 5) | count);
		return;
	}
	const uint32_t top = fs->hkey[0];
	if (price <= (top >> 5)) {
		const uint32_t en = top & 31u;
		fs->ent_pk[en] = pk;
		fs->ent_price[en] = price;
		heap_sink(fs, 0, (price << 5) | en, count);
	}
}

__device__ __forceinline__ uint32_t tree_price(SmemU16 probs, SmemU32 price, uint32_t base,
                                               uint32_t nbits, uint32_t value)
{
	uint32_t c = 0, node = 1;
	for (uint32_t i = nbits; i-- > 0;) {
		uint32_t bit = (value >> i) & 1;
		c += bit_price(probs, price, base + node, bit);
		node = (node << 1) | bit;
	}
	return c;
}

// src/lzma_packet_encoder.c:42-63 priced from the pre-state
__device__ __forceinline__ uint32_t length_price(SmemU16 probs, SmemU32 price, uint32_t base,
                                                 uint32_t len)
{
	uint32_t v = len - 2;
	if (v < 8) return bit_price(probs, price, base, 0) + tree_price(probs, price, base + LEN_LOW, 3, v);
	uint32_t c = bit_price(probs, price, base, 1);
	if (v < 16) return c + bit_price(probs, price, base + 1, 0) + tree_price(probs, price, base + LEN_MID, 3, v - 8);
	return c + bit_price(probs, price, base + 1, 1) + tree_price(probs, price, base + LEN_HIGH, 8, v - 16);
}

// Literal price at the model's position (src/lzma_packet_encoder.c:106-136)
__device__ __forceinline__ uint32_t literal_price(SmemU16 probs, SmemU32 price, uint32_t ctx,
                                                  uint32_t byte, uint32_t mbyte)
{
	uint32_t c = bit_price(probs, price, S_ISMATCH + ctx, 0);
	for (uint32_t d = 0; d < 8; d++) {
		uint32_t slot, bit;
		lit_event(d, byte, ctx >= 7, mbyte, slot, bit);
		c += bit_price(probs, price, slot, bit);
	}
	return c;
}

// Distance price without the pos-slot tree: reverse/align tree + direct bits
__device__ __forceinline__ uint32_t dist_tail_price(SmemU16 probs, SmemU32 price, const DistParts& d)
{
	uint32_t c = d.direct << 11, node = 1, v = d.low;
	for (uint32_t i = 0; i < d.rbits; i++) {
		uint32_t bit = v & 1;
		c += bit_price(probs, price, d.rbase + node, bit);
		node = (node << 1) | bit;
		v >>= 1;
	}
	return c;
}

// Four bytes at any address, from the two aligned words around it (the input is padded, see
// mg_ctx_create): match extension compares a word per step instead of a byte.
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* __restrict__ p)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
	return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
}

// Fills fs with the top-k of position m.pos and pops them (fs->pop_order, worst first).
// probs: the warp's live model (read only here).  Returns the candidate count kept;
// fs->candidates receives how many candidates were enumerated.  The model is taken by value so
// that the caller's copy stays in registers across this (deliberately out-of-line) call.
__device__ __noinline__ uint32_t warp_find(int lane, SmemU16 probs, SmemU32 price,
                                           SmemU32 recip, FindScratch* fs, const uint8_t* __restrict__ data,
                                           uint32_t n, const uint32_t* __restrict__ occ_start,
                                           const uint32_t* __restrict__ occ, const Model m, uint64_t excluded,
                                           uint32_t k, long long give_up_at = 0, FindLimits limits = FindLimits{0, 0})
{
	uint32_t candidates = 0, chunks = 0;
	// the length-price tables below share their memory with the walk's window mirrors: a window copy
	// (cp.async) still in flight must have landed before they are written
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncwarp();
	const uint32_t pos = m.pos, ctx = m.ctx;
	if (lane == 0) {
		fs->count = 0;
		fs->pops = 0;
	}
	// packet headers, src/lzma_packet_encoder.c:13-40
	const uint32_t p_match = bit_price(probs, price, S_ISMATCH + ctx, 1);
	const uint32_t hdr_match = p_match + bit_price(probs, price, S_ISREP + ctx, 0);
	const uint32_t hdr_rep = p_match + bit_price(probs, price, S_ISREP + ctx, 1);
	const uint32_t g0_0 = bit_price(probs, price, S_ISREPG0 + ctx, 0), g0_1 = bit_price(probs, price, S_ISREPG0 + ctx, 1);
	const uint32_t g1_0 = bit_price(probs, price, S_ISREPG1 + ctx, 0), g1_1 = bit_price(probs, price, S_ISREPG1 + ctx, 1);
	const uint32_t g2_0 = bit_price(probs, price, S_ISREPG2 + ctx, 0), g2_1 = bit_price(probs, price, S_ISREPG2 + ctx, 1);
	const uint32_t r0l_0 = bit_price(probs, price, S_ISREP0LONG + ctx, 0);
	const uint32_t r0l_1 = bit_price(probs, price, S_ISREP0LONG + ctx, 1);
	const uint32_t hdr_lrep0 = hdr_rep + g0_0 + r0l_1;
	const uint32_t hdr_lrep1 = hdr_rep + g0_1 + g1_0;
	const uint32_t hdr_lrep2 = hdr_rep + g0_1 + g1_1 + g2_0;
	const uint32_t hdr_lrep3 = hdr_rep + g0_1 + g1_1 + g2_1;

	const uint32_t byte = data[pos];
	const uint32_t rep_byte = pos > 0 ? data[pos - m.rep0 - 1] : 0;
	__syncwarp();
	if (lane == 0) {
		// packet_enumerator.c:60-66: LITERAL, then SHORT_REP when the rep0 byte matches
		heap_offer(fs, k, PK_LITERAL, literal_price(probs, price, ctx, byte, rep_byte), excluded);
		if (pos > 0 && byte == rep_byte) heap_offer(fs, k, PK_SHORT_REP, hdr_rep + g0_0 + r0l_0, excluded);
	}
	if (lane == 0) candidates += 1 + (pos > 0 && byte == rep_byte);
	__syncwarp();

	// substring_enumerator.c:87-88
	if (pos != 0 && pos != n - 1) {
		const uint32_t key = ((uint32_t)byte << 8) | data[pos + 1];
		uint32_t begin = occ_start[key];
		const uint32_t end = occ_start[key + 1];
		if (limits.window | limits.max_occ) {
			// the bucket is ascending: two binary searches bound the occurrences that may be offered
			auto first_at_least = [&](uint32_t value) {
				uint32_t lo = begin, hi = end;
				while (lo < hi) {
					const uint32_t mid = lo + (hi - lo) / 2;
					if (occ[mid] < value) lo = mid + 1;
					else hi = mid;
				}
				return lo;
			};
			const uint32_t upper = first_at_least(pos);  // occurrences at or after pos are never offered
			if (limits.window != 0 && pos > limits.window) begin = first_at_least(pos - limits.window);
			if (limits.max_occ != 0 && upper - begin > limits.max_occ) begin = upper - limits.max_occ;
		}
		const uint32_t max_len = n - pos < MAX_MATCH ? n - pos : MAX_MATCH;
		uint32_t have_len[2] = {1, 1};  // len_price filled for lengths 2..have_len
		uint32_t min_lp[2] = {0xffffffffu, 0xffffffffu};  // cheapest length price filled so far
		uint32_t slot_cached = 0xffffffffu, slot_price2 = 0, slot_price3 = 0, slot_price4 = 0, slot_price5 = 0;
		int32_t pm_prefix = INT32_MAX, pm_short = INT32_MAX;  // see the running-maximum filters below
		int32_t pm_t2 = INT32_MAX, pm_t3 = INT32_MAX, pm_t4 = INT32_MAX;
		uint32_t pm_root = 0xffffffffu, pm_have = 0;
		// align tree prices (src/lzma_packet_encoder.c:97-102): value i on lane i
		DistParts ap;
		ap.pslot = 14;
		ap.nlow = 4;
		ap.low = (uint32_t)lane & 15;
		ap.rbase = S_ALIGN;
		ap.rbits = 4;
		ap.direct = 0;
		const uint32_t align_tab = dist_tail_price(probs, price, ap);
		// One step ahead of the walk through the bucket: the next step's occurrence and the four bytes behind
		// its bigram are loaded while this step is priced (a step is one warp's serial work: nothing else hides
		// the two dependent L2 round trips, occurrence list -> input bytes).
		const uint32_t q2 = load_u32_unaligned(data + pos + 2);
		uint32_t o_next = begin + (uint32_t)lane < end ? occ[begin + (uint32_t)lane] : 0xffffffffu;
		uint32_t r2_next = o_next < pos ? load_u32_unaligned(data + o_next + 2) : 0u;
		for (uint32_t chunk = begin; chunk < end; chunk += 32) {
			const uint32_t o = o_next, r2 = r2_next;
			{
				const uint32_t idxn = chunk + 32u + (uint32_t)lane;
				o_next = idxn < end ? occ[idxn] : 0xffffffffu;
				r2_next = o_next < pos ? load_u32_unaligned(data + o_next + 2) : 0u;
			}
			const bool valid = o < pos;  // ascending bucket: stop at the first occurrence >= pos
			const uint32_t valid_mask = __ballot_sync(FULL, valid);
			if (valid_mask == 0) break;
			chunks++;
			// a clock-boxed step: a bucket of 100 k occurrences takes tens of milliseconds on its own -
			// the caller takes the proposal back and draws it again in the next launch (FIND_GAVE_UP)
			if (give_up_at != 0 && (chunks & 63u) == 0 && clock64() >= give_up_at) {
				__syncwarp();
				return FIND_GAVE_UP;
			}
			uint32_t L = 0;
			if (valid) {
				const uint32_t x2 = q2 ^ r2;
				if (x2) {
					L = 2 + (((uint32_t)__ffs((int)x2) - 1u) >> 3);
				} else {
					L = 6;
					const uint8_t* q = data + pos;
					const uint8_t* r = data + o;
					while (L < max_len) {
						const uint32_t x = load_u32_unaligned(q + L) ^ load_u32_unaligned(r + L);
						if (x) {
							L += ((uint32_t)__ffs((int)x) - 1u) >> 3;
							break;
						}
						L += 4;
					}
				}
				L = L < max_len ? L : max_len;
			}
			const uint32_t dist = pos - o - 1;
			uint32_t rep_mask = 0;
			if (valid) {
				rep_mask = (dist == m.rep0 ? 1u : 0u) | (dist == m.rep1 ? 2u : 0u) | (dist == m.rep2 ? 4u : 0u) |
				           (dist == m.rep3 ? 8u : 0u);
			}
			// extend the shared length price tables to the longest match of this step
			const uint32_t need0 = __reduce_max_sync(FULL, L);
			const uint32_t need1 = __reduce_max_sync(FULL, rep_mask ? L : 0u);
			if (need0 > have_len[0]) {
				uint32_t mn = 0xffffffffu;
				for (uint32_t l = have_len[0] + 1 + (uint32_t)lane; l <= need0; l += 32) {
					const uint32_t v = length_price(probs, price, S_LEN, l);
					fs->len_price[0][l - 2] = v;
					mn = min(mn, v);
				}
				min_lp[0] = min(min_lp[0], __reduce_min_sync(FULL, mn));
				have_len[0] = need0;
			}
			if (need1 > have_len[1]) {
				uint32_t mn = 0xffffffffu;
				for (uint32_t l = have_len[1] + 1 + (uint32_t)lane; l <= need1; l += 32) {
					const uint32_t v = length_price(probs, price, S_REPLEN, l);
					fs->len_price[1][l - 2] = v;
					mn = min(mn, v);
				}
				min_lp[1] = min(min_lp[1], __reduce_min_sync(FULL, mn));
				have_len[1] = need1;
			}
			__syncwarp();
			// per-occurrence distance prices: tail once, pos-slot tree per length context.  The model
			// does not change during a find and a bucket is walked in ascending position order, so a
			// lane sees long stretches of one pos slot: its four tree prices are kept in registers.
			const DistParts dp = dist_parts(valid ? dist : 0);
			const uint32_t align_price = __shfl_sync(FULL, align_tab, (int)(dp.low & 15));
			uint32_t base2 = 0, base3 = 0, base4 = 0, base5 = 0;
			if (valid) {
				uint32_t tail = hdr_match;
				if (dp.pslot >= 14)
					tail += (dp.direct << 11) + align_price;
				else
					tail += dist_tail_price(probs, price, dp);
				if (dp.pslot != slot_cached) {
					slot_cached = dp.pslot;
					slot_price2 = tree_price(probs, price, S_POSSLOT, 6, dp.pslot);
					slot_price3 = tree_price(probs, price, S_POSSLOT + 64, 6, dp.pslot);
					slot_price4 = tree_price(probs, price, S_POSSLOT + 128, 6, dp.pslot);
					slot_price5 = tree_price(probs, price, S_POSSLOT + 192, 6, dp.pslot);
				}
				base2 = tail + slot_price2;
				base3 = tail + slot_price3;
				base4 = tail + slot_price4;
				base5 = tail + slot_price5;
			}
			const uint32_t hdr_r0 = hdr_lrep0, hdr_r1 = hdr_lrep1, hdr_r2 = hdr_lrep2, hdr_r3 = hdr_lrep3;
			// cheapest candidate of this lane (lower bound test against the heap root)
			const bool full = fs->count >= k;
			const uint32_t root = full ? fs->hkey[0] >> 5 : 0xffffffffu;
			// Long matches (zero runs: every occurrence matches 273 bytes) would price 272 lengths each just to learn
			// that none beats the root.  floor((b + lp[len]) / len) <= root  <=>  b <= (root + 1) * len - 1 - lp[len], so a
			// running maximum of the right-hand side over the lengths answers "can ANY length of this occurrence enter
			// the heap" with one comparison.  It is kept in registers, nine lengths per lane (lane j: lengths up to
			// 10 + 9 j), looked up by shuffle; the bucket's upper end only makes it pass more, never less:
			// occurrences that pass are priced length by length as before.
			// For lengths up to 31 a second running maximum has one length per lane (lane j: lengths 5..j, where the
			// pos-slot context no longer depends on the length): with the thresholds of lengths 2, 3 and 4 it answers
			// the question EXACTLY, so only occurrences that do hold a candidate at or below the root are priced
			// length by length.
			if (MG_FINDER_PM && full && (root != pm_root || have_len[0] != pm_have)) {
				int32_t q = INT32_MIN, s5 = INT32_MIN;
				pm_t2 = pm_t3 = pm_t4 = INT32_MAX;
				if (root < 0x400000u) {
					for (uint32_t j = 0; j < 9; j++) {
						const uint32_t len = 2u + 9u * (uint32_t)lane + j;
						if (len <= have_len[0]) q = max(q, (int32_t)((root + 1u) * len - 1u) - (int32_t)fs->len_price[0][len - 2]);
					}
					if (lane >= 5 && (uint32_t)lane <= have_len[0])
						s5 = (int32_t)((root + 1u) * (uint32_t)lane - 1u) - (int32_t)fs->len_price[0][lane - 2];
					if (have_len[0] >= 2) pm_t2 = (int32_t)((root + 1u) * 2u - 1u) - (int32_t)fs->len_price[0][0];
					if (have_len[0] >= 3) pm_t3 = (int32_t)((root + 1u) * 3u - 1u) - (int32_t)fs->len_price[0][1];
					if (have_len[0] >= 4) pm_t4 = (int32_t)((root + 1u) * 4u - 1u) - (int32_t)fs->len_price[0][2];
				} else {
					q = INT32_MAX;
					s5 = lane >= 5 ? INT32_MAX : INT32_MIN;
				}
				for (int o = 1; o < 32; o <<= 1) {
					const int32_t t = __shfl_up_sync(FULL, q, o);
					const int32_t t5 = __shfl_up_sync(FULL, s5, o);
					if (lane >= o) {
						q = max(q, t);
						s5 = max(s5, t5);
					}
				}
				pm_prefix = q;
				pm_short = s5;
				pm_root = root;
				pm_have = have_len[0];
			}
			const int32_t pm_at = __shfl_sync(FULL, pm_prefix, valid ? (int)((L - 2u) / 9u) : 0);
			const int32_t pm_sh = __shfl_sync(FULL, pm_short, valid && L < 31u ? (int)L : 31);
			uint32_t cheapest = 0xffffffffu;
			uint32_t live_lens = 0;  // bit (len-2): some candidate of that length is at or below the root
			uint32_t first_live = 2; // shortest length with such a candidate (what a replay of a match longer than 33 starts at)
			if (valid) {
				uint32_t rep_hdr_best = 0xffffffffu;
				if (rep_mask & 1) rep_hdr_best = min(rep_hdr_best, hdr_r0);
				if (rep_mask & 2) rep_hdr_best = min(rep_hdr_best, hdr_r1);
				if (rep_mask & 4) rep_hdr_best = min(rep_hdr_best, hdr_r2);
				if (rep_mask & 8) rep_hdr_best = min(rep_hdr_best, hdr_r3);
				// No candidate of this occurrence is cheaper per byte than its cheapest header + distance
				// plus the cheapest length price, spread over its longest length.  Once the heap is full
				// that bound is above the root for nearly every occurrence, which then costs O(1)
				// instead of one pricing per length.
				uint32_t lo = base2;
				if (L > 2) lo = min(lo, base3);
				if (L > 3) lo = min(lo, base4);
				if (L > 4) lo = min(lo, base5);
				uint32_t bound = per_byte(lo + min_lp[0], L, recip);
				if (rep_mask) bound = min(bound, per_byte(rep_hdr_best + min_lp[1], L, recip));
				if (!full) {
					cheapest = 0;
					live_lens = 0xffffffffu;
				} else if (bound <= root &&
				           (!MG_FINDER_PM || rep_mask != 0 || pm_root != root ||
				            (L > 31u ? (int32_t)lo <= pm_at
				                     : ((int32_t)base2 <= pm_t2 || (L >= 3u && (int32_t)base3 <= pm_t3) || (L >= 4u && (int32_t)base4 <= pm_t4) ||
				                        (L >= 5u && (int32_t)base5 <= pm_sh))))) {
					first_live = 0xffffffffu;
					for (uint32_t len = 2; len <= L; len++) {
						const uint32_t b = len == 2 ? base2 : len == 3 ? base3 : len == 4 ? base4 : base5;
						const uint32_t pm = per_byte(b + fs->len_price[0][len - 2], len, recip);
						uint32_t pr = 0xffffffffu;
						if (rep_mask) pr = per_byte(rep_hdr_best + fs->len_price[1][len - 2], len, recip);
						cheapest = min(cheapest, min(pm, pr));
						// lengths worth replaying: a bit set while L <= 33, the first such length beyond (see the replay below)
						if (min(pm, pr) <= root) {
							live_lens |= 1u << ((len - 2) & 31);
							first_live = min(first_live, len);
						}
					}
				}
				candidates += (L - 1) * (1 + __popc(rep_mask));
			}
			uint32_t pass = __ballot_sync(FULL, valid && cheapest <= root);
			// the one candidate that is never offered (top_k_packet_finder.c:99-101) would break the
			// shortcuts below if it were the cheap one; a step that contains it takes the full replay
			const uint32_t ex_type = pk_type(excluded), ex_idx = pk_dist(excluded), ex_len = pk_len(excluded);
			const bool ex_here = valid && ex_len >= 2 && ex_len <= L &&
			                     ((ex_type == T_MATCH && ex_idx == dist) ||
			                      (ex_type == T_LONG_REP && ex_idx < 4 && ((rep_mask >> ex_idx) & 1u)));
			if (full && pass && !__any_sync(FULL, ex_here)) {
				// With a full heap a candidate whose price EQUALS the root's only overwrites the root
				// entry in place (strict comparisons in the sifts leave the heap order alone,
				// max_heap.c:93-103), so of a stretch of ties only the last one survives, and a
				// strictly cheaper candidate that follows evicts even that one.  Ties are the norm
				// (integer cost/len over mostly untrained length/distance models), so: lanes in front
				// of the first strictly cheaper lane are skipped, and a step holding nothing but ties
				// collapses to one store by its last tying lane.
				const uint32_t less = __ballot_sync(FULL, valid && cheapest < root);
				if (less == 0) {
					const int who = 31 - __clz((int)pass);
					uint32_t done = 0;
					if (lane == who) {
						// last candidate of this lane, in enumeration order, that ties and is offered
						uint64_t last = 0;
						for (uint32_t len = 2; len <= L; len++) {
							const uint32_t b = len == 2 ? base2 : len == 3 ? base3 : len == 4 ? base4 : base5;
							if (per_byte(b + fs->len_price[0][len - 2], len, recip) == root && pk_pack(T_MATCH, dist, len) != excluded)
								last = pk_pack(T_MATCH, dist, len);
							if (rep_mask) {
								const uint32_t lp = fs->len_price[1][len - 2];
								if ((rep_mask & 1) && per_byte(hdr_r0 + lp, len, recip) == root && pk_pack(T_LONG_REP, 0, len) != excluded) last = pk_pack(T_LONG_REP, 0, len);
								if ((rep_mask & 2) && per_byte(hdr_r1 + lp, len, recip) == root && pk_pack(T_LONG_REP, 1, len) != excluded) last = pk_pack(T_LONG_REP, 1, len);
								if ((rep_mask & 4) && per_byte(hdr_r2 + lp, len, recip) == root && pk_pack(T_LONG_REP, 2, len) != excluded) last = pk_pack(T_LONG_REP, 2, len);
								if ((rep_mask & 8) && per_byte(hdr_r3 + lp, len, recip) == root && pk_pack(T_LONG_REP, 3, len) != excluded) last = pk_pack(T_LONG_REP, 3, len);
							}
						}
						if (last != 0) {
							fs->ent_pk[fs->hkey[0] & 31u] = last;
							done = 1;
						}
					}
					// the only tie of that lane was the excluded packet: fall back to the full replay
					if (__shfl_sync(FULL, done, who)) pass = 0;
					__syncwarp();
				} else {
					pass &= ~((1u << (__ffs((int)less) - 1)) - 1u);
				}
			}
			// replay survivors in enumeration order: occurrence-major, length ascending,
			// MATCH before LONG_REP 0..3 (packet_enumerator.c:47-54)
			while (pass) {
				const int who = __ffs(pass) - 1;
				pass &= pass - 1;
				if (lane == who) {
					// The root only falls while this lane replays, so lengths that were above it when
					// the lane was priced can be skipped; heap_offer re-tests the rest against the
					// live root.  (The bit set covers lengths 2..33; a longer match replays from the first
					// length that was at or below the root: prices per byte mostly fall with the length, so
					// that is where the candidates worth offering start.)
					uint32_t todo = L <= 33 ? live_lens : 0xffffffffu;
					for (uint32_t len = L <= 33 ? 2u : first_live; len <= L; len++) {
						if (L <= 33) {
							if (todo == 0) break;
							len = 2 + (uint32_t)__ffs((int)todo) - 1;
							todo &= todo - 1;
						}
						const uint32_t b = len == 2 ? base2 : len == 3 ? base3 : len == 4 ? base4 : base5;
						heap_offer(fs, k, pk_pack(T_MATCH, dist, len), per_byte(b + fs->len_price[0][len - 2], len, recip),
						           excluded);
						if (rep_mask) {
							const uint32_t lp = fs->len_price[1][len - 2];
							if (rep_mask & 1) heap_offer(fs, k, pk_pack(T_LONG_REP, 0, len), per_byte(hdr_r0 + lp, len, recip), excluded);
							if (rep_mask & 2) heap_offer(fs, k, pk_pack(T_LONG_REP, 1, len), per_byte(hdr_r1 + lp, len, recip), excluded);
							if (rep_mask & 4) heap_offer(fs, k, pk_pack(T_LONG_REP, 2, len), per_byte(hdr_r2 + lp, len, recip), excluded);
							if (rep_mask & 8) heap_offer(fs, k, pk_pack(T_LONG_REP, 3, len), per_byte(hdr_r3 + lp, len, recip), excluded);
						}
					}
				}
				__syncwarp();
			}
			if (valid_mask != FULL) break;
		}
	}
	__syncwarp();
	candidates = __reduce_add_sync(FULL, candidates);
	// pop everything, worst first (top_k_packet_finder.c:127-138)
	if (lane == 0) {
		fs->candidates = candidates;
		fs->chunks = chunks;
		uint32_t pops = 0;
		uint32_t count = fs->count;
		while (count > 0) {
			fs->pop_order[pops++] = (uint8_t)(fs->hkey[0] & 31u);
			count--;
			if (count) heap_sink(fs, 0, fs->hkey[count], count);
		}
		fs->count = 0;
		fs->pops = pops;
	}
	__syncwarp();
	return fs->pops;
}

}  // namespace mg
