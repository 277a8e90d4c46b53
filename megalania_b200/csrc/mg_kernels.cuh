// Kernels of the annealing hot path (sm_100a).  One warp = one model = one chain / slab / query.
#pragma once
#include "mg_device.cuh"
#include "mg_finder.cuh"

namespace mg {

constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr uint32_t RECIP_ENTRIES = 288;

// Dynamic shared memory of one CTA.
struct WarpShared {
	Record rec;
	FindScratch fs;
	uint64_t bar;
};
struct CtaShared {
	uint16_t price[2048];
	uint32_t recip[RECIP_ENTRIES];
	WarpShared warp[WARPS_PER_CTA];
};
static_assert(sizeof(WarpShared) % 16 == 0, "per-warp shared block must keep the record 16-byte aligned");
static_assert(offsetof(CtaShared, warp) % 16 == 0, "record alignment");

struct Tables {
	const uint16_t* price;  // [2048] floor(-log2(i/2048)*2048), reference generate_table.py:7-9
	const uint32_t* recip;  // [RECIP_ENTRIES]
};

__device__ __forceinline__ void cta_tables_load(CtaShared* sh, const Tables& t)
{
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh->price[i] = t.price[i];
	for (int i = threadIdx.x; i < (int)RECIP_ENTRIES; i += blockDim.x) sh->recip[i] = t.recip[i];
	if ((threadIdx.x & 31) == 0) mbar_init(&sh->warp[threadIdx.x >> 5].bar, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
}

// Error bits a walk can raise
constexpr uint32_t ERR_BAD_PACKET = 1, ERR_NOT_BOUNDARY = 2, ERR_OUTPUT_FULL = 4;

// Memory-safety check of a packet about to be priced at model m (the reference asserts /
// reads out of bounds instead; src/lzma_packet_encoder.c:186-190).
__device__ __forceinline__ bool packet_ok(const Model& m, uint32_t n, uint32_t type, uint32_t len, uint32_t dist)
{
	if (len == 0 || len > n - m.pos) return false;
	switch (type) {
	case T_LITERAL: return len == 1;
	case T_SHORT_REP: return len == 1 && m.rep0 < m.pos;
	case T_MATCH: return len >= 2 && len <= MAX_MATCH && dist < m.pos;
	case T_LONG_REP: return len >= 2 && len <= MAX_MATCH && dist < 4 && model_rep(m, dist) < m.pos;
	default: return false;
	}
}

// Running totals of one walk
struct Tally {
	uint64_t total;  // uniform: cost flushed so far
	uint32_t acc;    // per lane: cost since the last flush
	uint32_t packets, bits;
};

__device__ __forceinline__ void tally_flush(Tally& t)
{
	t.total += __reduce_add_sync(FULL, t.acc);
	t.acc = 0;
}

// Prices slab packets from m.pos until `stop_pidx` packets or `stop_pos` bytes are reached.
// Writes a checkpoint each time the walk crosses a multiple of `stride` when ck != nullptr
// (ck indexed by slot-1, slot = pos / stride).
__device__ __forceinline__ uint32_t walk_plain(int lane, uint16_t* probs, const uint16_t* price, Record* rec, Model& m,
                                               Tally& t, Window& w, const uint64_t* __restrict__ slab,
                                               const uint8_t* __restrict__ data, uint32_t n, uint32_t stop_pos,
                                               uint32_t stop_pidx, Record* ck, uint32_t* ck_pos, uint32_t* ck_pidx,
                                               uint32_t stride)
{
	uint32_t next_ck = ck ? (m.pos / stride + 1) * stride : 0xffffffffu;
	while (m.pos < stop_pos && m.pidx < stop_pidx) {
		if ((m.pos & ~31u) != w.base) {
			tally_flush(t);
			window_seek(lane, w, slab, data, n, m.pos);
		}
		const uint64_t pk = window_packet(w, m.pos);
		const uint32_t byte = window_byte(w, m.pos);
		const uint32_t type = pk_type(pk), len = pk_len(pk), dist = pk_dist(pk);
		if (!packet_ok(m, n, type, len, dist)) return ERR_BAD_PACKET;
		uint32_t mbyte = 0;
		if (type == T_LITERAL && m.ctx >= 7) mbyte = data[m.pos - m.rep0 - 1];
		t.bits += apply_packet(lane, probs, price, m, type, len, dist, byte, mbyte, t.acc);
		t.packets++;
		if (m.pos >= next_ck && m.pos < n) {
			tally_flush(t);
			const uint32_t slot = m.pos / stride;
			record_store(lane, rec, m, t.total, ck + (slot - 1));
			if (lane == 0) {
				if (ck_pos) ck_pos[slot - 1] = m.pos;
				if (ck_pidx) ck_pidx[slot - 1] = m.pidx;
			}
			next_ck = (slot + 1) * stride;
		}
	}
	tally_flush(t);
	return 0;
}

// ---- K3a: score whole slabs ---------------------------------------------------------------------
struct ScoreArgs {
	const uint8_t* data;
	uint32_t n;
	const uint64_t* slabs;   // [nslabs][n] packed
	uint32_t nslabs;
	uint32_t stop_pos;       // price [0, stop_pos)
	uint64_t* out_cost;      // [nslabs]
	uint32_t* out_count;     // [nslabs] live packets, may be null
	uint32_t* out_err;       // [nslabs]
	Record* ck;              // [nslabs][nslots-1] or null
	uint32_t* ck_pos;        // [nslabs][nslots-1] or null
	uint32_t* ck_pidx;       // [nslabs][nslots-1] or null
	uint32_t stride, nslots;
	Record* final_model;     // [nslabs] model after the walk, or null
	Tables tables;
};

__global__ void __launch_bounds__(CTA_THREADS) score_kernel(ScoreArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	CtaShared* sh = reinterpret_cast<CtaShared*>(smem_raw);
	cta_tables_load(sh, a.tables);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t s = blockIdx.x * WARPS_PER_CTA + warp;
	if (s >= a.nslabs) return;
	WarpShared* ws = &sh->warp[warp];
	Model m;
	model_init(lane, ws->rec.probs, m);
	Tally t = {0, 0, 0, 0};
	Window w;
	w.base = 0xffffffffu;
	const size_t ckoff = (size_t)s * (a.nslots - 1);
	uint32_t err = walk_plain(lane, ws->rec.probs, sh->price, &ws->rec, m, t, w, a.slabs + (size_t)s * a.n, a.data, a.n,
	                          a.stop_pos, 0xffffffffu, a.ck ? a.ck + ckoff : nullptr,
	                          a.ck_pos ? a.ck_pos + ckoff : nullptr, a.ck_pidx ? a.ck_pidx + ckoff : nullptr, a.stride);
	if (!err && m.pos != a.stop_pos) err = ERR_NOT_BOUNDARY;
	if (a.final_model) record_store(lane, &ws->rec, m, t.total, a.final_model + s);
	if (lane == 0) {
		a.out_cost[s] = t.total;
		if (a.out_count) a.out_count[s] = m.pidx;
		a.out_err[s] = err;
	}
}

// ---- K2: top-k queries ---------------------------------------------------------------------------
struct TopkArgs {
	const uint8_t* data;
	uint32_t n;
	const uint32_t* occ_start;
	const uint32_t* occ;
	const uint64_t* slab;  // packed, n slots
	int state_mode;
	const uint32_t* positions;
	uint32_t npos, k;
	const Record* ck;  // mode 1: checkpoints of `slab`
	const uint32_t* ck_pos;
	uint32_t nslots, stride;
	uint64_t* out_pk;      // [npos][k]
	uint32_t* out_price;   // [npos][k]
	int32_t* out_count;    // [npos]
	uint32_t* out_err;     // [npos]
	unsigned long long* candidates;
	Tables tables;
};

__global__ void __launch_bounds__(CTA_THREADS) topk_kernel(TopkArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	CtaShared* sh = reinterpret_cast<CtaShared*>(smem_raw);
	cta_tables_load(sh, a.tables);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WarpShared* ws = &sh->warp[warp];
	uint32_t parity = 0;
	uint32_t cand = 0;
	for (uint32_t q = blockIdx.x * WARPS_PER_CTA + warp; q < a.npos; q += gridDim.x * WARPS_PER_CTA) {
		const uint32_t qpos = a.positions[q];
		Model m;
		uint32_t err = 0;
		if (qpos >= a.n) {
			err = ERR_NOT_BOUNDARY;
		} else if (a.state_mode == 0) {
			model_init(lane, ws->rec.probs, m);
			m.pos = qpos;
		} else {
			// last checkpoint at or before the query
			uint32_t below = 0;
			for (uint32_t j = lane; j + 1 < a.nslots; j += 32) below += a.ck_pos[j] <= qpos ? 1u : 0u;
			below = __reduce_add_sync(FULL, below);
			uint64_t cost = 0;
			if (below == 0)
				model_init(lane, ws->rec.probs, m);
			else
				record_load(lane, &ws->rec, m, cost, a.ck + (below - 1), &ws->bar, parity);
			Tally t = {cost, 0, 0, 0};
			Window w;
			w.base = 0xffffffffu;
			err = walk_plain(lane, ws->rec.probs, sh->price, &ws->rec, m, t, w, a.slab, a.data, a.n, qpos, 0xffffffffu,
			                 nullptr, nullptr, nullptr, a.stride);
			if (!err && m.pos != qpos) err = ERR_NOT_BOUNDARY;
		}
		uint32_t pops = 0;
		if (!err) {
			pops = warp_find(lane, ws->rec.probs, sh->price, sh->recip, &ws->fs, a.data, a.n, a.occ_start, a.occ, m,
			                 a.slab[qpos], a.k);
			cand += ws->fs.candidates;
			for (uint32_t i = lane; i < pops; i += 32) {
				const uint32_t e = ws->fs.pop_order[i];
				a.out_pk[(size_t)q * a.k + i] = ws->fs.ent_pk[e];
				a.out_price[(size_t)q * a.k + i] = ws->fs.ent_price[e];
			}
		}
		if (lane == 0) {
			a.out_count[q] = (int32_t)pops;
			a.out_err[q] = err;
		}
		__syncwarp();
	}
	if (lane == 0 && cand) atomicAdd(a.candidates, (unsigned long long)cand);
}

// ---- K3/K4: the annealing loop ---------------------------------------------------------------------
struct Edit {
	uint32_t pos;
	uint32_t pad;
	uint64_t pk;
};

struct ChainState {
	uint64_t rng;
	uint64_t cur_cost;
	uint64_t best_cost;
	uint32_t live_count;
	uint32_t err;
};

struct ChainStats {
	unsigned long long evals, attempts, accepted, new_best, packets, bits, slab_bytes, ck_bytes, finds, candidates,
	    edits, overflows;
};

struct TraceRec {
	uint64_t cost;
	uint32_t flags;
	uint32_t undo_count;
};

struct AnnealArgs {
	const uint8_t* data;
	uint32_t n;
	const uint32_t* occ_start;
	const uint32_t* occ;
	Tables tables;
	uint32_t chains, k, stride, nslots, log_cap, track_best;
	uint64_t* slabs;     // [chains][n]
	uint64_t* bests;     // [chains][n] or null
	Record* ck;          // [chains][2][nslots-1]
	uint32_t* ck_pidx;   // [chains][2][nslots-1]
	uint8_t* ck_live;    // [chains][nslots-1] which of the two buffers is current
	Edit* logs;          // [chains][log_cap]
	ChainState* state;   // [chains]
	ChainStats* stats;   // [chains]
	TraceRec* trace;     // [chains][trace_cap] or null
	uint32_t trace_cap;
	uint32_t* attempts_out;  // [chains]
	// run parameters
	uint32_t evals, max_attempts, schedule, step, num_iters, first_eval;
	const float* temps;
};

// memcmp(data+pos-d-1, data+pos, len) == 0 across the warp (src/packet_slab_neighbour.c:74-80)
__device__ __forceinline__ bool rep_matches(int lane, const uint8_t* __restrict__ data, uint32_t pos, uint32_t d,
                                            uint32_t len)
{
	bool same = true;
	const uint8_t* a = data + pos;
	const uint8_t* b = data + (pos - d - 1);
	for (uint32_t i = lane; i < len; i += 32) same = same && a[i] == b[i];
	return __all_sync(FULL, same);
}

struct EditLog {
	Edit* e;
	uint32_t cap;
	uint32_t stored;    // physical entries
	uint32_t count;     // logical edits (what the reference's undo stack would hold)
	uint32_t dup_pos;   // position whose entry may be rewritten (pos+1 of a shrink), or ~0
	uint32_t dup_index;
	bool overflow;
};

__device__ __forceinline__ void log_put(int lane, EditLog& lg, uint32_t pos, uint64_t pk)
{
	lg.count++;
	if (pos == lg.dup_pos) {
		if (lane == 0) lg.e[lg.dup_index].pk = pk;
		return;
	}
	if (lg.stored >= lg.cap) {
		lg.overflow = true;
		return;
	}
	if (lane == 0) {
		lg.e[lg.stored].pos = pos;
		lg.e[lg.stored].pk = pk;
	}
	lg.stored++;
}

// src/packet_slab_neighbour.c:48-72.  Returns false when the finder has no alternative.
__device__ __forceinline__ bool pick_from_topk(int lane, WarpShared* ws, const CtaShared* sh, const AnnealArgs& a,
                                               const Model& m, uint64_t excluded, bool best, uint64_t& rng,
                                               uint64_t& chosen, unsigned long long& cand)
{
	const uint32_t count =
	    warp_find(lane, ws->rec.probs, sh->price, sh->recip, &ws->fs, a.data, a.n, a.occ_start, a.occ, m, excluded, a.k);
	cand += ws->fs.candidates;
	if (count == 0) return false;
	uint32_t choice = rng31(rng) % count;
	for (int i = 1; i < 8; i++) {
		const uint32_t c = rng31(rng) % count;
		choice = c > choice ? c : choice;
	}
	if (rng31(rng) % 8 == 0 || best) choice = count - 1;
	chosen = ws->fs.ent_pk[ws->fs.pop_order[choice]];
	return true;
}

__global__ void __launch_bounds__(CTA_THREADS) anneal_kernel(AnnealArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	CtaShared* sh = reinterpret_cast<CtaShared*>(smem_raw);
	cta_tables_load(sh, a.tables);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t c = blockIdx.x * WARPS_PER_CTA + warp;
	if (c >= a.chains) return;
	WarpShared* ws = &sh->warp[warp];
	uint16_t* probs = ws->rec.probs;
	const uint16_t* price = sh->price;
	const uint8_t* __restrict__ data = a.data;
	const uint32_t n = a.n, nck = a.nslots - 1;

	uint64_t* slab = a.slabs + (size_t)c * n;
	Record* ck = a.ck + (size_t)c * 2 * nck;
	uint32_t* ck_pidx = a.ck_pidx + (size_t)c * 2 * nck;
	uint8_t* ck_live = a.ck_live + (size_t)c * nck;
	ChainState st = a.state[c];
	uint64_t rng = st.rng;
	uint32_t parity = 0;
	ChainStats cs = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	EditLog lg;
	lg.e = a.logs + (size_t)c * a.log_cap;
	lg.cap = a.log_cap;
	const float temp = a.temps ? a.temps[c] : 0.f;

	uint32_t done = 0, attempts = 0;
	while (done < a.evals && attempts < a.max_attempts && st.err == 0) {
		attempts++;
		lg.stored = lg.count = 0;
		lg.dup_pos = 0xffffffffu;
		lg.dup_index = 0;
		lg.overflow = false;

		// ---- pick the packet to mutate and the checkpoint to start from ------------------
		const uint32_t target = rng31(rng) % st.live_count;  // neighbour.c:163
		uint32_t below = 0;
		for (uint32_t j = lane; j < nck; j += 32) below += ck_pidx[(size_t)ck_live[j] * nck + j] <= target ? 1u : 0u;
		const uint32_t j0 = __reduce_add_sync(FULL, below);  // checkpoint slot to resume from (0 = start of file)
		Model m;
		Tally t = {0, 0, 0, 0};
		if (j0 == 0) {
			model_init(lane, probs, m);
		} else {
			record_load(lane, &ws->rec, m, t.total, ck + (size_t)ck_live[j0 - 1] * nck + (j0 - 1), &ws->bar, parity);
			cs.ck_bytes += sizeof(Record);
		}
		Window w;
		w.base = 0xffffffffu;
		const uint32_t start_pos = m.pos;

		// ---- prefix: neighbour.c:22-32 from the checkpoint instead of from byte 0 ----------
		uint32_t err = walk_plain(lane, probs, price, &ws->rec, m, t, w, slab, data, n, n, target, nullptr, nullptr,
		                          nullptr, a.stride);
		if (err || m.pos >= n) {
			st.err = err ? err : ERR_NOT_BOUNDARY;
			break;
		}

		// ---- mutate: neighbour.c:119-152 ---------------------------------------------------
		const uint32_t pos = m.pos;
		if ((pos & ~31u) != w.base) window_seek(lane, w, slab, data, n, pos);
		const uint64_t first = window_packet(w, pos);
		const uint32_t byte0 = window_byte(w, pos);
		uint64_t newpk = 0;
		uint64_t override_pk = 0;
		uint32_t override_pos = 0xffffffffu;
		bool mutated = false;
		if (pos + 1 < n && rng31(rng) % 2 == 0) {
			const uint64_t second = slab[pos + 1];
			const uint32_t ft = pk_type(first), flen = pk_len(first);
			const uint32_t stype = pk_type(second), slen = pk_len(second), sdist = pk_dist(second);
			if ((ft == T_LONG_REP || ft == T_MATCH) && flen > 2) {
				newpk = PK_LITERAL;
				override_pos = pos + 1;
				override_pk = pk_pack(ft, pk_dist(first), flen - 1);
				log_put(lane, lg, pos, newpk);
				lg.dup_index = lg.stored;
				log_put(lane, lg, pos + 1, override_pk);
				lg.dup_pos = pos + 1;  // a later repair of this slot rewrites the entry in place
				mutated = true;
			} else if ((ft == T_LITERAL || ft == T_SHORT_REP) && (stype == T_MATCH || stype == T_LONG_REP)) {
				const int64_t src = (int64_t)pos - (int64_t)(stype == T_LONG_REP ? model_rep(m, sdist & 3) : sdist);
				if (slen < MAX_MATCH && src > 0 && byte0 == data[src - 1]) {
					newpk = pk_pack(stype, sdist, slen + 1);
					log_put(lane, lg, pos, newpk);
					mutated = true;
				}
			}
		}
		bool ok = true;
		if (!mutated) {
			cs.finds++;
			ok = pick_from_topk(lane, ws, sh, a, m, first, false, rng, newpk, cs.candidates);
			if (ok) log_put(lane, lg, pos, newpk);
		}
		if (!ok) {
			// no alternative at this position: not an evaluation (main.c:81-84)
			if (a.trace && attempts <= a.trace_cap && lane == 0) {
				TraceRec r = {0, 0, 0};
				a.trace[(size_t)c * a.trace_cap + attempts - 1] = r;
			}
			cs.packets += t.packets;
			cs.bits += t.bits;
			cs.slab_bytes += (uint64_t)(m.pos - start_pos) * 9;
			continue;
		}

		// ---- price the mutated packet (neighbour.c:169) ------------------------------------
		{
			const uint32_t type = pk_type(newpk), len = pk_len(newpk), dist = pk_dist(newpk);
			if (!packet_ok(m, n, type, len, dist)) {
				st.err = ERR_BAD_PACKET;
				break;
			}
			uint32_t mbyte = 0;
			if (type == T_LITERAL && m.ctx >= 7) mbyte = data[m.pos - m.rep0 - 1];
			t.bits += apply_packet(lane, probs, price, m, type, len, dist, byte0, mbyte, t.acc);
			t.packets++;
		}

		// ---- repair + price the rest (neighbour.c:82-117), checkpointing as we go ------------
		uint32_t next_ck = (j0 + 1) * a.stride;
		uint32_t seen = 0;
		while (m.pos < n) {
			seen++;
			if ((m.pos & ~31u) != w.base) {
				tally_flush(t);
				window_seek(lane, w, slab, data, n, m.pos);
			}
			// a checkpoint belongs to the boundary reached before this packet is touched
			if (m.pos >= next_ck) {
				tally_flush(t);
				const uint32_t slot = m.pos / a.stride;
				const uint32_t buf = ck_live[slot - 1] ^ 1u;
				record_store(lane, &ws->rec, m, t.total, ck + (size_t)buf * nck + (slot - 1));
				if (lane == 0) ck_pidx[(size_t)buf * nck + (slot - 1)] = m.pidx;
				cs.ck_bytes += sizeof(Record);
				next_ck = (slot + 1) * a.stride;
			}
			const uint64_t old = m.pos == override_pos ? override_pk : window_packet(w, m.pos);
			const uint32_t byte = window_byte(w, m.pos);
			uint64_t pk = old;
			uint32_t type = pk_type(pk);
			const uint32_t rep_byte = data[m.pos - m.rep0 - 1];
			if (type == T_SHORT_REP || type == T_LITERAL) {
				if (byte == rep_byte) {
					if (seen < 4) pk = PK_SHORT_REP;
				} else {
					pk = PK_LITERAL;
				}
			} else if (type == T_LONG_REP) {
				const uint32_t len = pk_len(pk);
				uint32_t idx_now = pk_dist(pk);
				if (len > n - m.pos || idx_now > 3) {
					st.err = ERR_BAD_PACKET;
					break;
				}
				bool good = rep_matches(lane, data, m.pos, model_rep(m, idx_now), len);
				for (uint32_t idx = 0; idx < 4 && !good; idx++) {
					idx_now = idx;
					good = rep_matches(lane, data, m.pos, model_rep(m, idx_now), len);
				}
				pk = pk_pack(T_LONG_REP, idx_now, len);
				if (!good) {
					const bool best = rng31(rng) % 4 == 0;
					cs.finds++;
					uint64_t chosen = pk;
					pick_from_topk(lane, ws, sh, a, m, pk, best, rng, chosen, cs.candidates);
					pk = chosen;
				}
			}
			if (pk != old) log_put(lane, lg, m.pos, pk);
			type = pk_type(pk);
			const uint32_t len = pk_len(pk), dist = pk_dist(pk);
			if (!packet_ok(m, n, type, len, dist)) {
				st.err = ERR_BAD_PACKET;
				break;
			}
			t.bits += apply_packet(lane, probs, price, m, type, len, dist, byte, rep_byte, t.acc);
			t.packets++;
		}
		if (st.err) break;
		tally_flush(t);
		cs.packets += t.packets;
		cs.bits += t.bits;
		cs.slab_bytes += (uint64_t)(n - start_pos) * 9;
		cs.edits += lg.count;
		if (lg.overflow) {
			// accept/reject buffer too small for this proposal: drop it, uncounted
			cs.overflows++;
			if (a.trace && attempts <= a.trace_cap && lane == 0) {
				TraceRec r = {0, 0, lg.count};
				a.trace[(size_t)c * a.trace_cap + attempts - 1] = r;
			}
			continue;
		}

		// ---- accept / reject: main.c:86-96 ---------------------------------------------------
		const uint64_t cost = t.total;
		const uint32_t r = rng31(rng);
		bool uphill;
		if (a.schedule == 0) {
			const uint64_t i = (uint64_t)a.first_eval + done;
			const uint64_t mod = i * i + 1 + (uint64_t)a.step * (uint64_t)a.num_iters / 2;
			const uint64_t x = (uint64_t)r % mod;
			uphill = x * x < (uint64_t)a.num_iters;
		} else {
			const float e = -__logf(((float)r + 0.5f) * (1.0f / 2147483648.0f));
			uphill = (float)(cost - st.cur_cost) <= temp * e;
		}
		uint32_t flags = 1;
		if (st.cur_cost == 0 || cost < st.cur_cost || uphill) {
			flags |= 2;
			st.cur_cost = cost;
			st.live_count = m.pidx;
			// commit the accept/reject buffer and the checkpoints written on the way
			__syncwarp();
			for (uint32_t i = lane; i < lg.stored; i += 32) slab[lg.e[i].pos] = lg.e[i].pk;
			for (uint32_t j = j0 + lane; j < nck; j += 32) ck_live[j] ^= 1;
			__syncwarp();
			cs.accepted++;
			if (st.best_cost == 0 || cost < st.best_cost) {
				st.best_cost = cost;
				flags |= 4;
				cs.new_best++;
				if (a.track_best) {
					__threadfence_block();
					uint64_t* best = a.bests + (size_t)c * n;
					for (uint32_t i = lane; i < n; i += 32) best[i] = slab[i];
				}
			}
		}
		if (a.trace && attempts <= a.trace_cap && lane == 0) {
			TraceRec rec = {cost, flags, lg.count};
			a.trace[(size_t)c * a.trace_cap + attempts - 1] = rec;
		}
		done++;
		__syncwarp();
	}

	cs.evals = done;
	cs.attempts = attempts;
	if (lane == 0) {
		st.rng = rng;
		a.state[c] = st;
		a.stats[c] = cs;
		a.attempts_out[c] = attempts;
	}
}

// ---- K5: range coder over one slab (src/range_encoder.c) -------------------------------------------
struct EncodeArgs {
	const uint8_t* data;
	uint32_t n;
	const uint64_t* slab;
	uint8_t* out;
	uint32_t cap;
	uint32_t* out_len;
	uint32_t* out_err;
	Tables tables;
};

struct RangeCoder {
	uint64_t low;
	uint32_t range;
	uint32_t cache;
	uint64_t cache_size;
	uint8_t* out;
	uint32_t cap, len;
};

__device__ __forceinline__ void rc_put(RangeCoder& rc, uint32_t b)
{
	if (rc.len < rc.cap) rc.out[rc.len] = (uint8_t)b;
	rc.len++;
}

// src/range_encoder.c:18-38
__device__ __forceinline__ void rc_shift_low(RangeCoder& rc)
{
	const uint32_t hi = (uint32_t)(rc.low >> 32), lo = (uint32_t)rc.low;
	if (lo < 0xFF000000u || hi != 0) {
		uint32_t first = rc.cache;
		do {
			rc_put(rc, (first + (hi & 0xff)) & 0xff);
			first = 0xff;
		} while (--rc.cache_size != 0);
		rc.cache = (lo >> 24) & 0xff;
	}
	rc.cache_size++;
	rc.low = (uint64_t)(lo << 8);
}

// src/range_encoder.c:47-64
__device__ __forceinline__ void rc_bit(RangeCoder& rc, uint32_t bit, uint32_t prob)
{
	const uint32_t bound = (rc.range >> 11) * prob;
	if (bit) {
		rc.low += bound;
		rc.range -= bound;
	} else {
		rc.range = bound;
	}
	while ((rc.range & 0xFF000000u) == 0) {
		rc.range <<= 8;
		rc_shift_low(rc);
	}
}

// src/range_encoder.c:66-81
__device__ __forceinline__ void rc_direct(RangeCoder& rc, uint32_t bits, uint32_t nbits)
{
	while (nbits) {
		nbits--;
		rc.range >>= 1;
		if ((bits >> nbits) & 1) rc.low += rc.range;
		if ((rc.range & 0xFF000000u) == 0) {
			rc.range <<= 8;
			rc_shift_low(rc);
		}
	}
}

__global__ void __launch_bounds__(32) encode_kernel(EncodeArgs a)
{
	__shared__ __align__(16) Record rec;
	__shared__ uint32_t ev[32];
	const int lane = threadIdx.x;
	Model m;
	model_init(lane, rec.probs, m);
	RangeCoder rc = {0, 0xFFFFFFFFu, 0, 1, a.out, a.cap, 0};
	Window w;
	w.base = 0xffffffffu;
	uint32_t err = 0;
	while (m.pos < a.n) {
		window_seek(lane, w, a.slab, a.data, a.n, m.pos);
		const uint64_t pk = window_packet(w, m.pos);
		const uint32_t byte = window_byte(w, m.pos);
		const uint32_t type = pk_type(pk), len = pk_len(pk), dist = pk_dist(pk);
		if (!packet_ok(m, a.n, type, len, dist)) {
			err = ERR_BAD_PACKET;
			break;
		}
		uint32_t mbyte = 0;
		if (type == T_LITERAL && m.ctx >= 7) mbyte = a.data[m.pos - m.rep0 - 1];
		DistParts dp;
		dp.pslot = dp.nlow = dp.low = dp.rbase = dp.rbits = dp.direct = 0;
		if (type == T_MATCH) dp = dist_parts(dist);
		uint32_t slot = 0, bit = 0;
		const bool active = packet_event(lane, type, len, dist, m.ctx, byte, mbyte, dp, slot, bit);
		uint32_t word = 0;
		if (active) {
			uint32_t p = rec.probs[slot];
			word = (bit << 15) | p;
			p = bit ? p - (p >> 5) : p + ((2048u - p) >> 5);
			rec.probs[slot] = (uint16_t)p;
		}
		ev[lane] = word;
		uint32_t mask = __ballot_sync(FULL, active);
		__syncwarp();
		if (lane == 0) {
			// lanes are in coding order; the direct bits of a far match sit before lane 18
			const uint32_t direct_val = dp.direct ? (dist & ((1u << dp.nlow) - 1)) >> 4 : 0;
			bool direct_done = dp.direct == 0;
			while (mask) {
				const int l = __ffs(mask) - 1;
				mask &= mask - 1;
				if (!direct_done && l >= 18) {
					rc_direct(rc, direct_val, dp.direct);
					direct_done = true;
				}
				const uint32_t e = ev[l];
				rc_bit(rc, e >> 15, e & 0x7fff);
			}
			if (!direct_done) rc_direct(rc, direct_val, dp.direct);
		}
		__syncwarp();
		model_advance(m, type, len, dist);
	}
	if (lane == 0) {
		for (int i = 0; i < 5; i++) rc_shift_low(rc);  // src/range_encoder.c:40-45
		*a.out_len = rc.len;
		*a.out_err = err | (rc.len > rc.cap ? ERR_OUTPUT_FULL : 0);
	}
}

// ---- slab format conversion: host LZMAPacket (12 B) <-> packed u64 ----------------------------------
__global__ void pack_kernel(const uint32_t* __restrict__ raw, uint64_t* __restrict__ packed, size_t count)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t type = raw[i * 3] & 0xff, dist = raw[i * 3 + 1], len = raw[i * 3 + 2] & 0xffff;
		packed[i] = pk_pack(type, dist, len);
	}
}

__global__ void unpack_kernel(const uint64_t* __restrict__ packed, uint32_t* __restrict__ raw, size_t count)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
		const uint64_t p = packed[i];
		raw[i * 3] = pk_type(p);
		raw[i * 3 + 1] = pk_dist(p);
		raw[i * 3 + 2] = pk_len(p);
	}
}

__global__ void fill_literal_kernel(uint64_t* __restrict__ packed, size_t count)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		packed[i] = PK_LITERAL;
}

// Every slot of a slab a chain may walk must be decodable on its own terms
// (type, length, extent, absolute match distance).
__global__ void validate_kernel(const uint64_t* __restrict__ packed, uint32_t n, uint32_t* bad)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint64_t p = packed[i];
		const uint32_t type = pk_type(p), len = pk_len(p), dist = pk_dist(p);
		bool ok = len >= 1 && len <= n - i;
		if (type == T_LITERAL || type == T_SHORT_REP) ok = ok && len == 1;
		else if (type == T_MATCH) ok = ok && len >= 2 && len <= MAX_MATCH && dist < i;
		else if (type == T_LONG_REP) ok = ok && len >= 2 && len <= MAX_MATCH && dist < 4;
		else ok = false;
		if (!ok) atomicAdd(bad, 1u);
	}
}

// ---- K1: bigram index (src/substring_enumerator.c:26-47) ---------------------------------------------
// Stable counting sort of positions 0..n-2 by (data[i], data[i+1]).  Each warp owns a contiguous
// range of positions and a private histogram row, so the scatter needs no atomics and stays
// in ascending position order inside a bucket.
constexpr uint32_t INDEX_KEYS = 65536;

__device__ __forceinline__ void index_walk(const uint8_t* __restrict__ data, uint32_t n, uint32_t* row, uint32_t begin,
                                           uint32_t end, const uint32_t* start, uint32_t* occ)
{
	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1;
	for (uint32_t base = begin; base < end; base += 32) {
		const uint32_t i = base + lane;
		const bool in = i < end;
		const uint32_t key = in ? ((uint32_t)data[i] << 8 | data[i + 1]) : 0xffffffffu;
		const uint32_t peers = __match_any_sync(FULL, key);
		const uint32_t rank = __popc(peers & lt);
		uint32_t slot = 0;
		if (in && rank == 0) {
			slot = row[key];
			row[key] = slot + __popc(peers);
		}
		slot = __shfl_sync(FULL, slot, __ffs(peers) - 1);
		if (in && occ) occ[start[key] + slot + rank] = i;
		__syncwarp();
	}
	(void)n;
}

__global__ void index_count_kernel(const uint8_t* __restrict__ data, uint32_t n, uint32_t* rows, uint32_t per_warp)
{
	const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t begin = wid * per_warp;
	const uint32_t total = n - 1;
	if (begin >= total) return;
	const uint32_t end = begin + per_warp < total ? begin + per_warp : total;
	index_walk(data, n, rows + (size_t)wid * INDEX_KEYS, begin, end, nullptr, nullptr);
}

// rows[w][key] -> exclusive prefix over w; totals[key] = bucket size
__global__ void index_prefix_kernel(uint32_t* rows, uint32_t nwarps, uint32_t* totals)
{
	const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
	if (key >= INDEX_KEYS) return;
	uint32_t run = 0;
	for (uint32_t w = 0; w < nwarps; w++) {
		const uint32_t c = rows[(size_t)w * INDEX_KEYS + key];
		rows[(size_t)w * INDEX_KEYS + key] = run;
		run += c;
	}
	totals[key] = run;
}

// start[key] = exclusive scan of totals, start[65536] = n-1.  One block of 1024 threads.
__global__ void __launch_bounds__(1024) index_scan_kernel(const uint32_t* totals, uint32_t* start)
{
	__shared__ uint32_t part[1024];
	const uint32_t tid = threadIdx.x;
	uint32_t local[64];
	uint32_t sum = 0;
	for (int i = 0; i < 64; i++) {
		local[i] = sum;
		sum += totals[tid * 64 + i];
	}
	part[tid] = sum;
	__syncthreads();
	for (uint32_t off = 1; off < 1024; off <<= 1) {
		uint32_t v = tid >= off ? part[tid - off] : 0;
		__syncthreads();
		part[tid] += v;
		__syncthreads();
	}
	const uint32_t base = part[tid] - sum;
	for (int i = 0; i < 64; i++) start[tid * 64 + i] = base + local[i];
	if (tid == 1023) start[INDEX_KEYS] = part[1023];
}

__global__ void index_scatter_kernel(const uint8_t* __restrict__ data, uint32_t n, uint32_t* rows, uint32_t per_warp,
                                     const uint32_t* start, uint32_t* occ)
{
	const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t begin = wid * per_warp;
	const uint32_t total = n - 1;
	if (begin >= total) return;
	const uint32_t end = begin + per_warp < total ? begin + per_warp : total;
	index_walk(data, n, rows + (size_t)wid * INDEX_KEYS, begin, end, start, occ);
}

}  // namespace mg
