// Counter-addressable MT19937 (jump-ahead) + AS241 Brownian increment generation.
//
// Replaces BrownianMotionFromMersenneRandomNumbers.doGenerateBrownianMotion
// (J/montecarlo/BrownianMotionFromMersenneRandomNumbers.java:141-191): ONE sequential MersenneTwister(seed) stream, draw
// order path -> time -> factor, two 32-bit words per uniform.  Here the stream is cut into B contiguous sub-streams (one
// per thread block); the 624-word state at the head of each sub-stream is obtained by jump-ahead:
//
//   x[k+J] = sum_i g_i x[k+i]   with  g(x) = x^J mod phi(x),  phi = characteristic polynomial (degree 19937),
//
// i.e. a jumped state is a GF(2) linear combination of shifted copies of the source sequence (no sequential Horner loop:
// every output word is an independent XOR reduction, which is what the GPU wants).  phi is found once per process with
// Berlekamp-Massey on the generator's own output (no tables to download), jump polynomials by square-and-multiply.
// Sub-stream heads are filled by a doubling tree: level k applies g_{chunk*2^k} to heads [0,2^k) to get heads [2^k,2^(k+1)).
//
// HBM layout of the result: one slab double[T*F][paths] ("[t][f][path]"), each row a RandomVariable of the caller.
#include "fmb_common.cuh"
#include "fmb_icdf.cuh"
#include <algorithm>
#include <memory>

namespace fmb {

// ================================================================================================================
// Host: MT19937 seeding (commons-math3 MersenneTwister(long) == init_by_array{hi, lo}) and raw sequence
// ================================================================================================================
static const int MT_N = 624, MT_M = 397;
static const int DEG = 19937;

static void mtSeedState(int64_t seed, uint32_t* st) {
	const uint32_t key[2] = { (uint32_t)((uint64_t)seed >> 32), (uint32_t)((uint64_t)seed & 0xffffffffull) };
	st[0] = 19650218u;
	for (int i = 1; i < MT_N; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
	int i = 1, j = 0;
	for (int k = MT_N; k > 0; k--) {
		st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
		if (++i >= MT_N) { st[0] = st[MT_N - 1]; i = 1; }
		if (++j >= 2) j = 0;
	}
	for (int k = MT_N - 1; k > 0; k--) {
		st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
		if (++i >= MT_N) { st[0] = st[MT_N - 1]; i = 1; }
	}
	st[0] = 0x80000000u;
}

static inline uint32_t hostTwist(uint32_t a, uint32_t b) {
	const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
	return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// raw[k], k < n: raw[0..623] = state, raw[k] = raw[k-227] ^ twist(raw[k-624], raw[k-623])
static void mtRawSequence(const uint32_t* st, size_t n, std::vector<uint32_t>& raw) {
	raw.resize(n);
	for (size_t k = 0; k < n && k < (size_t)MT_N; k++) raw[k] = st[k];
	for (size_t k = MT_N; k < n; k++) raw[k] = raw[k - (MT_N - MT_M)] ^ hostTwist(raw[k - MT_N], raw[k - MT_N + 1]);
}

// ================================================================================================================
// Host: GF(2)[x] arithmetic modulo the characteristic polynomial
// ================================================================================================================
static const int PW = (DEG + 64) / 64;            // words holding bits 0..DEG (312)
struct CharPoly {
	std::vector<int> exps;                        // exponents with coefficient 1, ascending, last == DEG
	int gap = 0;                                  // DEG - second highest exponent
	std::vector<uint64_t> bits;                   // dense form, PW words
	bool ready = false;
};
static CharPoly g_phi;
static std::mutex g_phiMu;

// Berlekamp-Massey over GF(2) on one output bit-plane of the raw sequence.
static bool computeCharPoly() {
	uint32_t st[MT_N];
	mtSeedState(4357, st);
	const int NB = 2 * DEG + 64;
	std::vector<uint32_t> raw;
	mtRawSequence(st, (size_t)NB + 1, raw);
	std::vector<uint8_t> s(NB);
	for (int i = 0; i < NB; i++) s[i] = (uint8_t)(raw[i + 1] & 1u);   // skip raw[0] (its low 31 bits are not state)
	// reversed sequence in words so that a window of it lines up with the connection polynomial
	const int RW = (NB + 63) / 64 + 4;
	std::vector<uint64_t> R(RW, 0);
	for (int i = 0; i < NB; i++) if (s[i]) { const int j = NB - 1 - i; R[j >> 6] |= 1ull << (j & 63); }
	const int CWORDS = (DEG + 2 + 63) / 64 + 1;
	std::vector<uint64_t> Cc(CWORDS, 0), Bc(CWORDS, 0), Tc(CWORDS, 0);
	Cc[0] = 1; Bc[0] = 1;
	int L = 0, m = 1;
	for (int n = 0; n < NB; n++) {
		// d = sum_{i=0..L} C_i s[n-i];  s[n-i] = Rbit[NB-1-n+i]
		const int off = NB - 1 - n;
		const int w0 = off >> 6, sh = off & 63;
		uint64_t acc = 0;
		const int words = (L >> 6) + 1;
		for (int w = 0; w < words; w++) {
			uint64_t win = R[w0 + w] >> sh;
			if (sh) win |= R[w0 + w + 1] << (64 - sh);
			acc ^= win & Cc[w];
		}
		const int d = __builtin_parityll(acc);
		if (!d) { m++; continue; }
		const bool grow = (2 * L <= n);
		if (grow) Tc = Cc;
		// C ^= B << m
		const int ws = m >> 6, bs = m & 63;
		for (int w = CWORDS - 1; w >= ws; w--) {
			uint64_t v = Bc[w - ws] << bs;
			if (bs && w - ws - 1 >= 0) v |= Bc[w - ws - 1] >> (64 - bs);
			Cc[w] ^= v;
		}
		if (grow) { L = n + 1 - L; Bc = Tc; m = 1; } else m++;
	}
	if (L != DEG) return false;
	// characteristic polynomial: phi_{L-i} = C_i
	g_phi.bits.assign(PW, 0);
	g_phi.exps.clear();
	for (int i = L; i >= 0; i--) if ((Cc[i >> 6] >> (i & 63)) & 1ull) {
		const int e = L - i;
		g_phi.exps.push_back(e);
		g_phi.bits[e >> 6] |= 1ull << (e & 63);
	}
	if (g_phi.exps.back() != DEG || g_phi.exps.size() < 2) return false;
	g_phi.gap = DEG - g_phi.exps[g_phi.exps.size() - 2];
	// self-check: phi annihilates every bit-plane of the raw sequence (from index 1 on)
	for (int n = 1; n < 40; n++) {
		uint32_t acc = 0;
		for (int e : g_phi.exps) acc ^= raw[n + e];
		if (acc != 0) return false;
	}
	g_phi.ready = true;
	return true;
}

static int ensureCharPoly() {
	std::lock_guard<std::mutex> lk(g_phiMu);
	if (g_phi.ready) return FMB_OK;
	if (!computeCharPoly()) { setError("MT19937 characteristic polynomial derivation failed its self-check"); return FMB_ECUDA; }
	return FMB_OK;
}

typedef std::vector<uint64_t> Poly;               // PW words, degree < DEG

// v: 2*PW words holding a polynomial of degree < 2*DEG; reduce modulo phi in place; result in the low PW words
static void reduceModPhi(std::vector<uint64_t>& v) {
	const CharPoly& phi = g_phi;
	const int top = (int)v.size() - 1;
	if (phi.gap >= 64) {
		const int wDeg = DEG >> 6, bDeg = DEG & 63;
		for (int w = top; w >= wDeg; w--) {
			uint64_t chunk = v[w];
			if (w == wDeg) chunk &= ~((1ull << bDeg) - 1ull);
			if (!chunk) continue;
			v[w] ^= chunk;
			const long base = (long)w * 64 - DEG;
			for (size_t k = 0; k + 1 < phi.exps.size(); k++) {
				const long Tpos = base + phi.exps[k];
				if (Tpos >= 0) {
					const long tw = Tpos >> 6; const int ts = (int)(Tpos & 63);
					v[tw] ^= chunk << ts;
					if (ts) v[tw + 1] ^= chunk >> (64 - ts);
				} else {
					v[0] ^= chunk >> (-Tpos);
				}
			}
		}
	} else {
		for (int i = (int)v.size() * 64 - 1; i >= DEG; i--) {
			if (!((v[i >> 6] >> (i & 63)) & 1ull)) continue;
			for (int e : phi.exps) { const int t = i - DEG + e; v[t >> 6] ^= 1ull << (t & 63); }
		}
	}
}

static inline uint64_t spread32(uint32_t x) {     // interleave zeros: bit i -> bit 2i
	uint64_t v = x;
	v = (v | (v << 16)) & 0x0000ffff0000ffffull;
	v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
	v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
	v = (v | (v << 2)) & 0x3333333333333333ull;
	v = (v | (v << 1)) & 0x5555555555555555ull;
	return v;
}

static void polySquare(Poly& p) {
	std::vector<uint64_t> v(2 * PW + 1, 0);
	for (int w = 0; w < PW; w++) {
		v[2 * w] = spread32((uint32_t)(p[w] & 0xffffffffull));
		v[2 * w + 1] = spread32((uint32_t)(p[w] >> 32));
	}
	reduceModPhi(v);
	for (int w = 0; w < PW; w++) p[w] = v[w];
}

static void polyMulX(Poly& p) {
	uint64_t carry = 0;
	for (int w = 0; w < PW; w++) { const uint64_t nc = p[w] >> 63; p[w] = (p[w] << 1) | carry; carry = nc; }
	if ((p[DEG >> 6] >> (DEG & 63)) & 1ull) for (int w = 0; w < PW; w++) p[w] ^= g_phi.bits[w];
}

// a * b mod phi (shift-and-add over the set bits of b)
static Poly polyMul(const Poly& a, const Poly& b) {
	std::vector<uint64_t> v(2 * PW + 1, 0);
	for (int w = 0; w < PW; w++) {
		uint64_t bits = b[w];
		while (bits) {
			const int s = __builtin_ctzll(bits);
			bits &= bits - 1;
			if (s == 0) { for (int i = 0; i < PW; i++) v[w + i] ^= a[i]; }
			else { for (int i = 0; i < PW; i++) { v[w + i] ^= a[i] << s; v[w + i + 1] ^= a[i] >> (64 - s); } }
		}
	}
	reduceModPhi(v);
	return Poly(v.begin(), v.begin() + PW);
}

// x^J mod phi
static Poly polyPowX(uint64_t J) {
	Poly r(PW, 0);
	r[0] = 1;
	if (J == 0) return r;
	int top = 63 - __builtin_clzll(J);
	for (int b = top; b >= 0; b--) {
		polySquare(r);
		if ((J >> b) & 1ull) polyMulX(r);
	}
	return r;
}

// ascending exponents of the set coefficients, as 16-bit indices (degree < 19937), padded with JUMP_PAD to a multiple of 8.
// JUMP_PAD points at a zero word behind the expanded sequence, so padded entries XOR in nothing.
static const int JUMP_LIST_MAX = 19944;           // 19937 rounded up to a multiple of 8
static int polyToBitList(const Poly& p, uint16_t* out /* JUMP_LIST_MAX */, uint16_t pad) {
	int n = 0;
	for (int i = 0; i < DEG; i++) if ((p[i >> 6] >> (i & 63)) & 1ull) out[n++] = (uint16_t)i;
	const int padded = (n + 7) & ~7;
	for (int i = n; i < padded; i++) out[i] = pad;
	return padded;
}

static void polyToWords32(const Poly& p, uint32_t* out /* MT_N words */) {
	for (int i = 0; i < MT_N; i++) {
		const int w = i >> 1;
		out[i] = (w < PW) ? (uint32_t)((i & 1) ? (p[w] >> 32) : (p[w] & 0xffffffffull)) : 0u;
	}
}

// ================================================================================================================
// Device kernels
// ================================================================================================================
static const int SEQ_LEN = DEG + MT_N;            // 20561 raw words cover every x[n+i], n < 624, i < 19937
static const int JUMP_THREADS = 640;
static const int JUMP_PAD = SEQ_LEN;              // list padding entry: seq[n + JUMP_PAD] is a zero word for every n < 624
static const int JUMP_SMEM_WORDS = SEQ_LEN + MT_N;

// One level of the radix-8 jump tree.  grid (numOutputs, segments).  Output o = (j - 1) * have + i is source state i (< have) jumped with
// polynomial j (1..7) of the level: dst[o] = g_j applied to src[i].  Block (o, s) expands its source state to the raw words its share of
// the coefficient list needs (shared memory), then thread n XORs seq[n + e] over the set coefficients e in list_j[s*perSeg, (s+1)*perSeg).
// The list (ascending exponents, one per set coefficient, built once per polynomial on the host) replaces a bit-scan loop: eight
// independent shared-memory loads per iteration, addresses known up front.
static const int JUMP_RADIX = 8;
struct JumpLevel { int listLen[JUMP_RADIX - 1]; };
__global__ void __launch_bounds__(JUMP_THREADS) mtJumpApplyKernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int have,
		const uint16_t* __restrict__ lists, JumpLevel level, int perSeg, int atomicCombine) {
	extern __shared__ uint32_t seq[];
	const int tid = threadIdx.x;
	const int jm1 = blockIdx.x / have, i0 = blockIdx.x - jm1 * have;
	const uint32_t* s = src + (size_t)i0 * MT_N;
	const uint16_t* list = lists + (size_t)jm1 * JUMP_LIST_MAX;
	const int listLen = level.listLen[jm1];
	const int kBeg = blockIdx.y * perSeg;
	const int kEnd = min(listLen, kBeg + perSeg);
	if (kBeg >= kEnd) return;
	for (int i = tid; i < MT_N; i += JUMP_THREADS) { seq[i] = s[i]; seq[SEQ_LEN + i] = 0u; }
	__syncthreads();
	// the list is ascending: the last real entry of the segment bounds the raw words needed
	int last = kEnd - 1;
	while (last > kBeg && __ldg(list + last) == JUMP_PAD) last--;
	const int needEnd = min(SEQ_LEN, (int)__ldg(list + last) + MT_N);
	for (int j0 = MT_N; j0 < needEnd; j0 += (MT_N - MT_M)) {
		const int j = j0 + tid;
		if (tid < (MT_N - MT_M) && j < needEnd) seq[j] = seq[j - (MT_N - MT_M)] ^ mtTwist(seq[j - MT_N], seq[j - MT_N + 1]);
		__syncthreads();
	}
	if (tid < MT_N) {
		uint32_t acc0 = 0, acc1 = 0;
		const uint32_t* base = seq + tid;
#pragma unroll 4
		for (int k = kBeg; k < kEnd; k += 8) {                 // perSeg and listLen are multiples of 8, the list is 16-byte aligned
			const uint4 q = __ldg(reinterpret_cast<const uint4*>(list + k));
			acc0 ^= base[q.x & 0xffffu]; acc1 ^= base[q.x >> 16];
			acc0 ^= base[q.y & 0xffffu]; acc1 ^= base[q.y >> 16];
			acc0 ^= base[q.z & 0xffffu]; acc1 ^= base[q.z >> 16];
			acc0 ^= base[q.w & 0xffffu]; acc1 ^= base[q.w >> 16];
		}
		uint32_t* d = dst + (size_t)blockIdx.x * MT_N + tid;
		if (atomicCombine) atomicXor(d, acc0 ^ acc1); else *d = acc0 ^ acc1;
	}
}

// tempered outputs of one state (test hook fmb_mt_words): single block, sequential 227-wide refresh
__global__ void __launch_bounds__(256) mtWordsKernel(const uint32_t* __restrict__ state, uint32_t* __restrict__ out, uint64_t n) {
	__shared__ uint32_t ring[2048];
	const int tid = threadIdx.x;
	for (int i = tid; i < MT_N; i += 256) ring[i] = state[i];
	__syncthreads();
	uint64_t genEnd = MT_N;
	for (uint64_t o0 = 0; o0 < n; o0 += 227) {
		if (tid < 227) {
			const uint64_t j = genEnd + tid;
			ring[j & 2047] = ring[(j - 227) & 2047] ^ mtTwist(ring[(j - MT_N) & 2047], ring[(j - MT_N + 1) & 2047]);
		}
		__syncthreads();
		if (tid < 227 && o0 + tid < n) out[o0 + tid] = mtTemper(ring[(genEnd + tid) & 2047]);
		genEnd += 227;
		__syncthreads();
	}
}

__global__ void icdfKernel(const double* __restrict__ p, double* __restrict__ out, uint64_t n) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = inverseCumulativeNormal(p[i]);
}

__global__ void uniformsFromWordsKernel(const uint32_t* __restrict__ w, double* __restrict__ out, uint64_t n) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = mtUniform(w[2 * i], w[2 * i + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// Brownian increments.  Block b owns paths [b*ppb, min(P,(b+1)*ppb)) and the MT sub-stream that starts at its first word.
// The raw MT19937 words live in a shared-memory ring of eight 624-word blocks.  (1) One block is always being generated ahead of the
// consumers (227 threads, three words each, every index a compile-time offset from the block base; completion is signalled through an
// mbarrier that is waited for one batch later).  (2) Each thread turns two tempered words into a uniform, applies AS241 (central
// rational for every lane, straight-line; tail draws are parked in a per-warp queue and evaluated by full warps), scales by sqrt(dt)
// and drops the value into a shared-memory tile laid out [c = t*F+f][path in tile], (3) full tiles are written to HBM as rows of
// consecutive paths (one bulk store per row, or coalesced element-wise stores for narrow tiles).
// Algorithmic HBM bytes: 8 per increment (write only).
// ---------------------------------------------------------------------------------------------------------------
static const int BM_THREADS = 320;               // 10 warps: 312 uniforms per 624-word block keep 97.5 % of the lanes busy; two CTAs per SM
static const int BM_THREADS_WIDE = 640;          // used when the tile is so wide (large T*F) that only one CTA per SM fits
static const int RING_BLOCKS = 8;                // see the static_assert in bmGenerateKernel
static const int RING = RING_BLOCKS * MT_N;      // raw words
static const int RING_ALLOC = RING + 2;          // (+ padding: thread 169 reads one word past the last block before replacing the value)
static const int BM_HEADER = 16;                 // bytes in front of the ring (tail-queue counter)

// Tail draws (|u - 0.5| > 0.425, 15 % of all) cost ~3x a central draw (log, sqrt, a second rational) and would make
// almost every warp execute both branches.  They are parked in a shared-memory queue and evaluated by full warps.  Every WARP owns a
// slice of the queue and keeps its fill count in a register (ballot + popc give each lane its position): no atomics, no shuffles, no
// block-wide barrier around a drain.  A drain evaluates the dense multiple-of-32 prefix and moves the remainder (< 32) to the front.
template <bool SCALED> __device__ __forceinline__ uint32_t bmDrainWarp(double* __restrict__ qP, uint32_t* __restrict__ qSlot, uint32_t count, bool all,
		double* __restrict__ tile, int lane, float invPad, const double* __restrict__ sqrtDtPerColumn) {
	const uint32_t full = all ? count : (count & ~31u);
	for (uint32_t i = lane; i < full; i += 32) {
		const double p = qP[i];
		const uint32_t slot = qSlot[i];
		double v = as241Tail(p, p - 0.5);
		if (SCALED) {
			// column of the slot = slot / nPad; slot < 2^22, so (slot + 0.5) / nPad truncated in single precision is exact
			const uint32_t c = __float2uint_rz(((float)slot + 0.5f) * invPad);
			v = v * __ldg(sqrtDtPerColumn + c);
		}
		tile[slot] = v;
	}
	const uint32_t rem = count - full;
	if (rem) {                                                     // (full >= 32 > rem: source and destination do not overlap)
		if ((uint32_t)lane < rem) { qP[lane] = qP[full + lane]; qSlot[lane] = qSlot[full + lane]; }
	}
	__syncwarp();
	return rem;
}

// ---- TMA (bulk asynchronous copy engine) helpers: a finished tile row leaves shared memory as ONE bulk store instead of n 8-byte
//      stores with per-element address arithmetic.  The generic-proxy writes of the tile are made visible to the async proxy with
//      fence.proxy.async before the barrier that precedes the stores.
__device__ __forceinline__ void fenceProxyAsyncShared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulkStoreRow(double* gdst, const double* ssrc, uint32_t bytes) {
	const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommitAndWaitRead() {
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// next 624 raw words: block nb from block nb-1 (mod RING_BLOCKS).  x[k+624] = x[k+397] ^ twist(x[k], x[k+1]).
// Thread i < 227 produces words i, i+227 and i+454 one after the other: word i+227 needs word i and word i+454 needs word i+227
// (its own results, kept in registers), everything else comes from the previous block - so the whole block needs ONE barrier.
// The only cross-thread input, x[624] = new word 0 for the very last word, is recomputed by that thread.
__device__ __forceinline__ void bmRefreshBlock(uint32_t* __restrict__ ring, uint32_t prevOff, uint32_t newOff, int tid) {
	// prevOff / newOff: word offsets of the previous and of the new block in the ring (loop-carried by the caller: every address below is
	// one register plus a compile-time offset)
	constexpr int W = MT_N - MT_M;                                 // 227
	if (tid < W) {
		uint32_t* nw = ring + newOff + tid;
		const uint32_t* od = ring + prevOff + tid;
		const uint32_t v0 = od[MT_M] ^ mtTwist(od[0], od[1]);
		nw[0] = v0;
		const uint32_t v1 = v0 ^ mtTwist(od[W], od[W + 1]);
		nw[W] = v1;
		if (tid < MT_N - 2 * W) {                                  // 170 words
			uint32_t next;
			if (tid == MT_N - 2 * W - 1) {                             // word 623 needs x[624], the new word 0 (thread 0 is writing it): recomputed
				const uint32_t* o0 = ring + prevOff;
				next = o0[MT_M] ^ mtTwist(o0[0], o0[1]);
			} else {
				next = od[2 * W + 1];
			}
			nw[2 * W] = v1 ^ mtTwist(od[2 * W], next);
		}
	}
}

// The refresh is decoupled from its consumers with an mbarrier instead of __syncthreads(): a warp ARRIVES when its part of block k+1 is
// written and only WAITS for block k+1 one batch later, after it has consumed a batch from the blocks before it - by then every warp
// has long arrived, so nobody stalls at the barrier (with __syncthreads() the single hottest instruction of the kernel was the
// branch behind the barrier: 11 % of all stall samples, profiles/r02_notes.md).  One refresh is always in flight.
__device__ __forceinline__ uint32_t bmSmemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bmBarInit(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bmSmemAddr(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bmBarArriveWarp(uint64_t* bar, int lane) {
	__syncwarp();                                                  // the warp's writes are ordered before lane 0's (releasing) arrive
	if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bmSmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bmBarWait(uint64_t* bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bmSmemAddr(bar)), "r"(parity) : "memory");
	} while (!done);
}

// TMA: the tile holds the increments already scaled by sqrt(dt) (row stride nPad even: 16-byte aligned rows) and every row is written with
// one bulk store; needs P, tileN even.  !TMA: standard normals in the tile (odd row stride), scaled and stored element by element.
// UNIFORM: the tile receives the uniforms themselves (IndependentIncrementsFromICDF: the caller applies its own inverse distribution
// functions); the host then passes a scale table of ones.
template <int NT, bool TMA, bool UNIFORM> __global__ void __launch_bounds__(NT, NT == BM_THREADS ? 2 : 1) bmGenerateKernel(const uint32_t* __restrict__ states, double* __restrict__ out,
		uint64_t P, uint64_t Pr, uint32_t TF, uint32_t colChunks, uint32_t ppb, uint32_t tileN, uint32_t nPad, uint32_t qCap,
		const double* __restrict__ sqrtDtPerColumn) {
	// P paths of TF columns each, row stride of the output Pr (= P).  colChunks > 1 (T*F too large for one path to fit the tile): the
	// "paths" are column chunks of the real paths - virtual path v = (real path v / colChunks, chunk v % colChunks), TF columns each, in
	// the same stream order - and Pr is the real number of paths; tiles then hold a single virtual path.
	extern __shared__ __align__(16) unsigned char smemRaw[];
	uint32_t* ring = reinterpret_cast<uint32_t*>(smemRaw + BM_HEADER);
	double* qP = reinterpret_cast<double*>(smemRaw + BM_HEADER + RING_ALLOC * sizeof(uint32_t));
	uint32_t* qSlot = reinterpret_cast<uint32_t*>(qP + qCap);          // qCap entries (even): tail draws parked until a dense drain
	// the tile starts 16-byte aligned (bulk stores read whole 16-byte units); an OFFSET from the shared-memory base, so that the compiler
	// keeps the shared address space (32-bit addresses, STS) - rounding the pointer itself made every tile access a generic 64-bit one.
	// !TMA: standard normals, sqrt(dt) of the column is applied when the tile is written out; TMA: already scaled
	const uint32_t tileOff = (uint32_t)(BM_HEADER + RING_ALLOC * sizeof(uint32_t) + (size_t)qCap * (sizeof(double) + sizeof(uint32_t)) + 15) & ~15u;
	double* tile = reinterpret_cast<double*>(smemRaw + tileOff);
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	// ring depth: when a thread starts writing a new block, at most 2*NT - 2 + 624 complete words are unconsumed and one more block
	// has just been completed; the slowest warp may still be reading the batch before (2*NT words further back: a warp passes the
	// wait for refresh k only after every warp has arrived for it, i.e. has finished the batch before the one in which k was issued)
	static_assert(NT >= MT_N - MT_M && 4 * NT + 3 * MT_N <= RING, "block size against the refresh width / ring depth");
	static_assert((BM_HEADER + RING_ALLOC * sizeof(uint32_t)) % 8 == 0, "queue alignment");
	// this warp's slice of the tail queue (a multiple of 32 entries, >= 64) and its fill count
	const uint32_t wCap = (qCap / (NT / 32)) & ~31u;
	double* wqP = qP + (uint32_t)warp * wCap;
	uint32_t* wqSlot = qSlot + (uint32_t)warp * wCap;
	uint32_t wcount = 0;
	const uint32_t ltMask = (1u << lane) - 1u;

	uint64_t* refreshBar = reinterpret_cast<uint64_t*>(smemRaw);      // (the 16-byte header)
	for (int i = tid; i < MT_N; i += NT) ring[i] = states[(size_t)blockIdx.x * MT_N + i];
	if (tid == 0) bmBarInit(refreshBar, NT / 32);
	__syncthreads();

	// ring position of the next unconsumed raw word (always even), number of complete but unconsumed words, word offsets of the newest
	// complete block and of the block in flight, parity of the refresh in flight
	uint32_t cpos = MT_N, prevOff = 0, newOff = MT_N, refreshParity = 0;
	int avail = 0;
	bmRefreshBlock(ring, prevOff, newOff, tid);
	bmBarArriveWarp(refreshBar, lane);
	const uint64_t pBeg = (uint64_t)blockIdx.x * ppb;
	const uint64_t pEnd = min(P, pBeg + (uint64_t)ppb);
	// (path in tile, column) of this thread's draw, advanced by NT draws per iteration without divisions
	const uint32_t stepP = NT / TF, stepC = NT % TF;
	const float invPad = 1.0f / (float)nPad;

	for (uint64_t p0 = pBeg; p0 < pEnd; p0 += tileN) {
		const uint32_t n = (uint32_t)min((uint64_t)tileN, pEnd - p0);
		const uint32_t U = n * TF;
		uint32_t pl = (uint32_t)tid / TF, c = (uint32_t)tid % TF;
		for (uint32_t u0 = 0; u0 < U; u0 += NT) {
			const uint32_t need = min((uint32_t)NT, U - u0);
			while (avail < (int)(2 * need)) {
				bmBarWait(refreshBar, refreshParity);                  // the block in flight is complete ...
				refreshParity ^= 1u;
				avail += MT_N;
				prevOff = newOff;
				newOff = (newOff == RING - MT_N) ? 0 : newOff + MT_N;
				bmRefreshBlock(ring, prevOff, newOff, tid);            // ... and the next one starts from it
				bmBarArriveWarp(refreshBar, lane);
			}
			const bool act = (uint32_t)tid < need;
			const uint32_t slot = c * nPad + pl;
			const double sdt = (TMA && !UNIFORM) ? __ldg(sqrtDtPerColumn + c) : 1.0;
			uint32_t j = cpos + 2 * tid;                               // even, pair never wraps (RING is even)
			if (j >= RING) j -= RING;
			// (a thread without a draw reads words that are not generated yet: harmless, always inside the ring)
			const uint2 ww = *reinterpret_cast<const uint2*>(ring + j);
			const double u = mtUniform(mtTemper(ww.x), mtTemper(ww.y));
			if (UNIFORM) {
				if (act) tile[slot] = u;
			} else {
				// the central rational is evaluated for every lane (practically every warp holds central draws, so a branch around it
				// saves nothing) with a division that has no out-of-line slow path: straight-line code
				const double q = u - 0.5;
				const bool isTail = act && !(fabs(q) <= 0.425);
				double v = as241Central(q);
				if (TMA) v = v * sdt;
				if (act && !isTail) tile[slot] = v;
				// (the whole warp is here: the batch loop is uniform across the block)
				const uint32_t tails = __ballot_sync(0xffffffffu, isTail);
				if (isTail) {
					const uint32_t pos = wcount + __popc(tails & ltMask);
					wqP[pos] = u;
					wqSlot[pos] = slot;
				}
				wcount += __popc(tails);
				// as soon as a full warp's worth of tails is parked, evaluate it: small, frequent drains keep the warps of the block in step
				// (a warp that drains six groups at once falls five batches behind and the others wait for it at the refresh barrier)
				if (wcount >= 32u) {
					__syncwarp();
					wcount = bmDrainWarp<TMA>(wqP, wqSlot, wcount, false, tile, lane, invPad, sqrtDtPerColumn);
				}
			}
			pl += stepP; c += stepC;
			if (c >= TF) { c -= TF; pl++; }
			cpos += 2 * need;
			if (cpos >= RING) cpos -= RING;
			avail -= (int)(2 * need);
		}
		if (!UNIFORM) {
			__syncwarp();
			wcount = bmDrainWarp<TMA>(wqP, wqSlot, wcount, true, tile, lane, invPad, sqrtDtPerColumn);
		}
		if (TMA) fenceProxyAsyncShared();
		__syncthreads();
		if (TMA) {
			// one bulk store per column: n consecutive paths of column cc, 8n bytes (n even), source row and destination both 16-byte aligned
			bool issued = false;
			for (uint32_t cc = tid; cc < TF; cc += NT) { bulkStoreRow(out + (size_t)cc * P + p0, tile + (size_t)cc * nPad, n * 8u); issued = true; }
			if (issued) bulkCommitAndWaitRead();                  // the tile may be overwritten once the engine has READ it
		} else if (n >= 32) {
			double* dst = out + (size_t)warp * P + p0 + lane;
			const double* srcRow = tile + warp * nPad + lane;
			for (uint32_t cc = warp; cc < TF; cc += NT / 32) {
				const double sdt = __ldg(sqrtDtPerColumn + cc);
#pragma unroll 4
				for (uint32_t i = 0; i + lane < n; i += 32) dst[i] = srcRow[i] * sdt;
				dst += (size_t)(NT / 32) * P;
				srcRow += (NT / 32) * nPad;
			}
		} else {
			// narrow tile (wide T*F): n consecutive paths per column, (column, path) advanced without a division per element
			uint32_t cc = (uint32_t)tid / n, i = (uint32_t)tid - cc * n;
			const uint32_t sC = NT / n, sI = NT - sC * n;
			double* obase = out + p0;
			const double* sdt = sqrtDtPerColumn;
			if (colChunks > 1) {                               // (n == 1) columns [kc*TF, (kc+1)*TF) of real path pr
				const uint64_t pr = p0 / colChunks;
				const uint32_t kc = (uint32_t)(p0 - pr * colChunks);
				obase = out + (size_t)kc * TF * Pr + pr;
				sdt += (size_t)kc * TF;
			}
			for (uint32_t idx = tid; idx < U; idx += NT) {
				obase[(size_t)cc * Pr + i] = tile[cc * nPad + i] * __ldg(sdt + cc);
				cc += sC; i += sI;
				if (i >= n) { i -= n; cc++; }
			}
		}
		__syncthreads();
	}
}

// ================================================================================================================
// Host orchestration
// ================================================================================================================
// Jump polynomials of one sub-stream length ("chunk", in words), radix 8: level k, j = 1..7: g = x^(j * 8^k * chunk) mod phi, as
// coefficient lists (JUMP_LIST_MAX uint16 each) on the device.  Built on demand (a level's j-th polynomial is the (j-1)-th times the
// first; the next level's first is the previous level's first raised to the 8th power: three squarings) and cached per chunk.
// Nine dependent launches of ~60 us (binary tree, B = 296) become three.
struct JumpLevelPolys {
	uint16_t* dev = nullptr;                       // (JUMP_RADIX - 1) lists
	int count = 0;                                 // polynomials built so far (j = 1..count)
	int listLen[JUMP_RADIX - 1] = {0};
	Poly first, last;                              // host copies of g_1 and of g_count
};
struct DevicePolys { std::vector<JumpLevelPolys> levels; };
static std::map<uint64_t, DevicePolys> g_polyCache;     // key: chunk (words)
static std::mutex g_polyMu;
static const int MAX_LEVELS = 20;                       // 8^20 sub-streams

// level `level` of chunk with at least `need` polynomials (1..7)
static int getLevelPolys(uint64_t chunk, int level, int need, const JumpLevelPolys** out) {
	FMB_TRY(ensureCharPoly());
	std::lock_guard<std::mutex> lk(g_polyMu);
	if (level >= MAX_LEVELS || need < 1 || need > JUMP_RADIX - 1) { setError("too many jump levels"); return FMB_EINVAL; }
	DevicePolys& dp = g_polyCache[chunk];
	dp.levels.reserve(MAX_LEVELS);                              // (callers keep a pointer to a level: no reallocation later)
	if ((int)dp.levels.size() <= level) dp.levels.resize(level + 1);
	for (int k = 0; k <= level; k++) {
		JumpLevelPolys& L = dp.levels[k];
		const int want = (k == level) ? need : 1;                   // lower levels: only their first polynomial is needed to climb
		if (!L.dev && (k == level)) FMB_CUDA(cudaMalloc(&L.dev, (size_t)(JUMP_RADIX - 1) * JUMP_LIST_MAX * sizeof(uint16_t)));
		while (L.count < want) {
			if (L.count == 0) {
				if (k == 0) L.first = polyPowX(chunk);
				else { L.first = dp.levels[k - 1].first; for (int q = 0; q < 3; q++) polySquare(L.first); }
				L.last = L.first;
			} else {
				L.last = polyMul(L.last, L.first);
			}
			if (!L.dev) FMB_CUDA(cudaMalloc(&L.dev, (size_t)(JUMP_RADIX - 1) * JUMP_LIST_MAX * sizeof(uint16_t)));
			std::vector<uint16_t> list(JUMP_LIST_MAX);
			L.listLen[L.count] = polyToBitList(L.last, list.data(), (uint16_t)JUMP_PAD);
			FMB_CUDA(cudaMemcpyAsync(L.dev + (size_t)L.count * JUMP_LIST_MAX, list.data(), JUMP_LIST_MAX * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx().stream));
			FMB_CUDA(cudaStreamSynchronize(ctx().stream));
			L.count++;
		}
	}
	*out = &dp.levels[level];
	return FMB_OK;
}

// dst[(j - 1) * have + i] = g_j applied to src[i], for the first `count` outputs
static int launchJump(const uint32_t* src, uint32_t* dst, int have, const JumpLevelPolys& L, int count) {
	static bool attrSet = false;
	const size_t smem = (size_t)JUMP_SMEM_WORDS * sizeof(uint32_t);
	if (!attrSet) {
		FMB_CUDA(cudaFuncSetAttribute(mtJumpApplyKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attrSet = true;
	}
	const int polys = (count + have - 1) / have;
	JumpLevel lv;
	int maxLen = 0;
	for (int j = 0; j < JUMP_RADIX - 1; j++) { lv.listLen[j] = j < polys ? L.listLen[j] : 0; maxLen = std::max(maxLen, lv.listLen[j]); }
	if (maxLen == 0) {                                         // zero polynomial cannot occur (x^J mod phi != 0); keep the output defined
		FMB_CUDA(cudaMemsetAsync(dst, 0, (size_t)count * MT_N * sizeof(uint32_t), ctx().stream));
		return FMB_OK;
	}
	// split the coefficient lists over several blocks while the level has fewer outputs than the machine has SM slots
	const int slots = 2 * ctx().smCount;
	// (about two waves of blocks: a level whose output count is not a multiple of the SM count would otherwise wait for a ragged last wave)
	int segs = std::max(1, std::min(32, (2 * slots + count - 1) / count));
	int perSeg = (((maxLen + segs - 1) / segs) + 7) & ~7;
	segs = (maxLen + perSeg - 1) / perSeg;
	if (segs > 1) FMB_CUDA(cudaMemsetAsync(dst, 0, (size_t)count * MT_N * sizeof(uint32_t), ctx().stream));
	mtJumpApplyKernel<<<dim3(count, segs), JUMP_THREADS, smem, ctx().stream>>>(src, dst, have, L.dev, lv, perSeg, segs > 1 ? 1 : 0);
	countLaunch();
	FMB_CUDA(cudaGetLastError());
	return FMB_OK;
}

// heads[0] <- state of MersenneTwister(seed) jumped by firstWord; heads[b] <- heads[0] jumped by b*chunk, b < B
static int buildStreamHeads(int64_t seed, uint64_t firstWord, uint64_t chunk, int B, uint32_t* heads /* device, B*MT_N */) {
	Context& c = ctx();
	uint32_t st[MT_N];
	mtSeedState(seed, st);
	if (firstWord == 0) {
		FMB_CUDA(cudaMemcpyAsync(heads, st, sizeof(st), cudaMemcpyHostToDevice, c.stream));
		FMB_CUDA(cudaStreamSynchronize(c.stream));
	} else {
		const JumpLevelPolys* L;
		FMB_TRY(getLevelPolys(firstWord, 0, 1, &L));
		void* tmp;
		FMB_TRY(poolAlloc(sizeof(st), &tmp));
		FMB_CUDA(cudaMemcpyAsync(tmp, st, sizeof(st), cudaMemcpyHostToDevice, c.stream));
		FMB_CUDA(cudaStreamSynchronize(c.stream));
		int rc = launchJump((const uint32_t*)tmp, heads, 1, *L, 1);
		poolFree(tmp, sizeof(st));
		FMB_TRY(rc);
	}
	// radix-8 tree: level k turns heads [0, 8^k) into heads [8^k, 8^(k+1)): head[j * 8^k + i] = head[i] jumped by j * 8^k * chunk
	int level = 0;
	for (int64_t have = 1; have < B; have *= JUMP_RADIX, level++) {
		const int cnt = (int)std::min<int64_t>((JUMP_RADIX - 1) * have, B - have);
		const JumpLevelPolys* L;
		FMB_TRY(getLevelPolys(chunk, level, (int)((cnt + have - 1) / have), &L));
		FMB_TRY(launchJump(heads, heads + (size_t)have * MT_N, (int)have, *L, cnt));
	}
	return FMB_OK;
}

} // namespace fmb

using namespace fmb;

extern "C" {

// Host-only self-test hook (no GPU needed): the jumped state computed with the same polynomial code the device path uses,
// applied on the CPU.  tests/ check it against the sequentially skipped oracle stream.
int fmb_test_host_jump(int64_t seed, uint64_t J, uint32_t* state_out /* 624 */, int* phi_weight, int* phi_gap) {
	FMB_TRY(ensureCharPoly());
	if (phi_weight) *phi_weight = (int)g_phi.exps.size();
	if (phi_gap) *phi_gap = g_phi.gap;
	uint32_t st[MT_N];
	mtSeedState(seed, st);
	std::vector<uint32_t> raw;
	mtRawSequence(st, SEQ_LEN, raw);
	Poly g = polyPowX(J);
	if (J >= 3) {
		// self-check of the general product the radix-8 jump tree is built from: x^a * x^(J-a) = x^J (mod phi), and a cube by squaring
		// and multiplying
		const uint64_t a = J / 3;
		const Poly pa = polyPowX(a);
		Poly sq = pa;
		polySquare(sq);
		if (polyMul(pa, polyPowX(J - a)) != g || polyMul(sq, pa) != polyPowX(3 * a)) { setError("test_host_jump: polynomial product mismatch"); return FMB_ECUDA; }
	}
	uint32_t w32[MT_N];
	polyToWords32(g, w32);
	for (int n = 0; n < MT_N; n++) {
		uint32_t acc = 0;
		for (int w = 0; w < MT_N; w++) {
			uint32_t bits = w32[w];
			while (bits) { const int b = __builtin_ctz(bits); bits &= bits - 1; acc ^= raw[n + 32 * w + b]; }
		}
		state_out[n] = acc;
	}
	return FMB_OK;
}

int fmb_mt_words(int64_t seed, uint64_t word_offset, uint64_t n, uint32_t* host_out) {
	FMB_TRY(requireInit());
	if (n == 0) return FMB_OK;
	if (!host_out) { setError("null output"); return FMB_EINVAL; }
	void* head; void* dout;
	FMB_TRY(poolAlloc(MT_N * sizeof(uint32_t), &head));
	int rc = poolAlloc(n * sizeof(uint32_t), &dout);
	if (rc) { poolFree(head, MT_N * sizeof(uint32_t)); return rc; }
	rc = buildStreamHeads(seed, word_offset, 0, 1, (uint32_t*)head);
	if (rc == FMB_OK) {
		mtWordsKernel<<<1, 256, 0, ctx().stream>>>((const uint32_t*)head, (uint32_t*)dout, n);
		countLaunch();
		cudaError_t e = cudaMemcpyAsync(host_out, dout, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx().stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
		if (e != cudaSuccess) { setError("mt_words: %s", cudaGetErrorString(e)); rc = FMB_ECUDA; }
	}
	poolFree(head, MT_N * sizeof(uint32_t));
	poolFree(dout, n * sizeof(uint32_t));
	return rc;
}

int fmb_mt_uniforms(int64_t seed, uint64_t uniform_offset, uint64_t n, double* host_out) {
	FMB_TRY(requireInit());
	if (n == 0) return FMB_OK;
	if (!host_out) { setError("null output"); return FMB_EINVAL; }
	void* head; void* dw; void* du;
	FMB_TRY(poolAlloc(MT_N * sizeof(uint32_t), &head));
	FMB_TRY(poolAlloc(2 * n * sizeof(uint32_t), &dw));
	FMB_TRY(poolAlloc(n * sizeof(double), &du));
	int rc = buildStreamHeads(seed, 2 * uniform_offset, 0, 1, (uint32_t*)head);
	if (rc == FMB_OK) {
		mtWordsKernel<<<1, 256, 0, ctx().stream>>>((const uint32_t*)head, (uint32_t*)dw, 2 * n);
		uniformsFromWordsKernel<<<gridFor(n, 256), 256, 0, ctx().stream>>>((const uint32_t*)dw, (double*)du, n);
		countLaunch(2);
		cudaError_t e = cudaMemcpyAsync(host_out, du, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
		if (e != cudaSuccess) { setError("mt_uniforms: %s", cudaGetErrorString(e)); rc = FMB_ECUDA; }
	}
	poolFree(head, MT_N * sizeof(uint32_t));
	poolFree(dw, 2 * n * sizeof(uint32_t));
	poolFree(du, n * sizeof(double));
	return rc;
}

int fmb_icdf(const double* host_p, uint64_t n, double* host_out) {
	FMB_TRY(requireInit());
	if (n == 0) return FMB_OK;
	void* dp; void* dq;
	FMB_TRY(poolAlloc(n * sizeof(double), &dp));
	FMB_TRY(poolAlloc(n * sizeof(double), &dq));
	cudaError_t e = cudaMemcpyAsync(dp, host_p, n * sizeof(double), cudaMemcpyHostToDevice, ctx().stream);
	icdfKernel<<<gridFor(n, 256), 256, 0, ctx().stream>>>((const double*)dp, (double*)dq, n);
	countLaunch();
	if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, dq, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
	poolFree(dp, n * sizeof(double));
	poolFree(dq, n * sizeof(double));
	if (e != cudaSuccess) { setError("icdf: %s", cudaGetErrorString(e)); return FMB_ECUDA; }
	return FMB_OK;
}

// uniform == false: Brownian increments ICDF(u) * sqrt_dt[t];  uniform == true: the uniforms u themselves (sqrt_dt ignored)
static int generateIncrements(int64_t seed, int T, int F, uint64_t paths, uint64_t path_offset, const double* sqrt_dt, bool uniform, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (T <= 0 || F <= 0 || (!sqrt_dt && !uniform) || !out) { setError("bm_generate: bad argument"); return FMB_EINVAL; }
	Context& c = ctx();
	const uint64_t TF = (uint64_t)T * F;
	if (paths == 0) {                                           // a rank that owns no paths of the logical simulation: zero-length increments
		Slab* empty = nullptr;
		FMB_TRY(newSlab(8, &empty));
		for (uint64_t i = 0; i < TF; i++) out[i] = newView(empty, (double*)empty->base, 0);
		return FMB_OK;
	}

	// shared memory: header + ring + tail queue + tile [TF][nPad].  Two blocks per SM when a >= 4-path tile fits in half of the 227 KB -
	// with the full tail queue if possible, else with a half-size one (wide T*F: two 320-thread blocks hide each other's barriers, one
	// 640-thread block cannot) - else one block with the whole of it.  (Three blocks per SM with smaller tiles measured slower: one more
	// level of jump-ahead heads costs more than the extra warps give, profiles/r01_notes.md.)
	// TMA flush (one bulk store per tile row) needs 16-byte aligned rows on both sides: an even number of paths and an even tile
	// width / row stride.  FMB_BM_TMA=0 forces the element-wise flush (A/B measurements, profiles/r02_notes.md).
	bool tma = (paths % 2 == 0);
	if (const char* e = getenv("FMB_BM_TMA")) tma = tma && atoi(e) != 0;
	// the tail queue: 64 entries per warp (31 parked + 32 new at most)
	uint32_t tileN = 0, qCap = (BM_THREADS / 32) * 64;
	size_t fixed = 0;
	int blocksPerSm = 1;
	// blocks per SM to try, most first: FMB_BM_CTAS=3 adds three 320-thread blocks per SM (A/B switch: more resident warps, but half
	// again as many sub-stream heads for the jump-ahead)
	int maxCtas = 2;
	if (const char* e = getenv("FMB_BM_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 3) maxCtas = v; }
	for (int ctas = maxCtas; ctas >= 1; ctas--) {
		const size_t budget = ctas == 3 ? 74 * 1024 : (ctas == 2 ? 112 * 1024 : 224 * 1024);
		qCap = ((ctas == 1 ? BM_THREADS_WIDE : BM_THREADS) / 32) * 64;
		fixed = BM_HEADER + RING_ALLOC * sizeof(uint32_t) + (size_t)qCap * (sizeof(double) + sizeof(uint32_t));
		fixed = (fixed + 15) & ~(size_t)15;                           // the tile starts 16-byte aligned
		const size_t avail = budget > fixed ? budget - fixed : 0;
		const uint64_t maxPad = avail / (TF * sizeof(double));
		blocksPerSm = ctas;
		// rows of at least 16 paths: bulk-store flush (row stride tileN + 2); narrower tiles (T*F in the thousands): element-wise flush
		// with the odd row stride tileN + 1 - bulk copies of 32-64 bytes do not pay (measured, profiles/r02_notes.md)
		if (tma && maxPad >= 18) { tileN = (uint32_t)std::min<uint64_t>(((maxPad - 2) / 4) * 4, 256); break; }
		if (maxPad >= 5) { tileN = (uint32_t)std::min<uint64_t>(((maxPad - 1) / 4) * 4, 256); tma = false; break; }
		if (ctas == 1) { tileN = maxPad >= 2 ? (uint32_t)(maxPad - 1) : (uint32_t)maxPad; tma = false; }
	}
	// T*F so large that not even one path fits the tile: cut every path into colChunks chunks of TFk columns (a divisor of T*F that fits) and
	// run the kernel on these "virtual paths" - the stream order is unchanged, only the flush addresses differ
	uint64_t TFk = TF, colChunks = 1;
	if (tileN == 0) {
		const uint64_t cap = (224 * 1024 - fixed) / sizeof(double);
		for (uint64_t d = 2; d <= TF / 512 && colChunks == 1; d++) if (TF % d == 0 && TF / d <= cap) colChunks = d;
		if (colChunks == 1) { setError("bm_generate: T*F = %llu has no divisor between 512 and %llu to cut the paths into tile-sized column chunks", (unsigned long long)TF, (unsigned long long)cap); return FMB_EUNSUPPORTED; }
		TFk = TF / colChunks;
		tileN = 1;
		tma = false;
	}
	if (tileN % 2 || tileN < 16) tma = false;
	// row stride: TMA: even, and = 2 (mod 8) doubles so that consecutive columns start 4 banks apart (a warp writes 32 consecutive
	// columns of one path: 4-way instead of 16-way conflicts); element-wise flush: odd (conflict-free column writes)
	uint32_t nPad = tileN >= 2 ? (tileN | 1u) : tileN;
	if (tma) nPad = tileN + 2;
	const size_t smem = fixed + (size_t)TFk * nPad * sizeof(double);
	const uint64_t vpaths = paths * colChunks;                    // (virtual) paths the kernel iterates over

	// sub-streams: enough blocks to fill the machine, but at least ~32k uniforms each so that jump-ahead stays a small fraction
	uint64_t Bmax = (uint64_t)c.smCount * blocksPerSm;            // one wave of equal sub-streams
	const uint64_t totalUniforms = paths * TF;
	Bmax = std::max<uint64_t>(1, std::min<uint64_t>(Bmax, totalUniforms / 32768 + 1));
	uint64_t ppb = (vpaths + Bmax - 1) / Bmax;
	ppb = ((ppb + tileN - 1) / tileN) * tileN;
	const int B = (int)((vpaths + ppb - 1) / ppb);
	if (ppb > 0xffffffffull) { setError("bm_generate: too many paths per block"); return FMB_EUNSUPPORTED; }
	const uint64_t chunk = ppb * 2ull * TFk;

	void* heads = nullptr;
	FMB_TRY(poolAlloc((size_t)B * MT_N * sizeof(uint32_t), &heads));
	int rc = buildStreamHeads(seed, path_offset * 2ull * TF, chunk, B, (uint32_t*)heads);

	Slab* slab = nullptr;
	void* dsq = nullptr;
	if (rc == FMB_OK) rc = poolAlloc(TF * sizeof(double), &dsq);
	if (rc == FMB_OK) {
		std::lock_guard<std::mutex> lk(c.scratchMu);
		rc = ensureScratch(TF * sizeof(double), 0);
		if (rc == FMB_OK) {
			double* h = (double*)c.pinned;
			for (int t = 0; t < T; t++) for (int f = 0; f < F; f++) h[(size_t)t * F + f] = uniform ? 1.0 : sqrt_dt[t];
			cudaError_t e = cudaMemcpyAsync(dsq, h, TF * sizeof(double), cudaMemcpyHostToDevice, c.stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
			if (e != cudaSuccess) { setError("bm_generate: %s", cudaGetErrorString(e)); rc = FMB_ECUDA; }
		}
	}
	if (rc == FMB_OK) rc = newSlab(TF * paths * sizeof(double), &slab);
	// one CTA per SM only (wide tile): 640 threads, so that the SM still has 20 resident warps
	const bool wide = blocksPerSm == 1;
	if (rc == FMB_OK) {
		auto launch = [&](auto kernel, int NT, int slot) -> int {
			static size_t attrSmem[8] = {0};
			if (smem > attrSmem[slot]) {
				FMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				attrSmem[slot] = smem;
			}
			kernel<<<B, NT, smem, c.stream>>>((const uint32_t*)heads, (double*)slab->base, vpaths, paths, (uint32_t)TFk, (uint32_t)colChunks, (uint32_t)ppb, tileN, nPad, qCap,
			                                  (const double*)dsq);
			countLaunch();
			FMB_CUDA(cudaGetLastError());
			return FMB_OK;
		};
		if (uniform) {
			if (wide) rc = tma ? launch(bmGenerateKernel<BM_THREADS_WIDE, true, true>, BM_THREADS_WIDE, 4) : launch(bmGenerateKernel<BM_THREADS_WIDE, false, true>, BM_THREADS_WIDE, 5);
			else rc = tma ? launch(bmGenerateKernel<BM_THREADS, true, true>, BM_THREADS, 6) : launch(bmGenerateKernel<BM_THREADS, false, true>, BM_THREADS, 7);
		} else {
			if (wide) rc = tma ? launch(bmGenerateKernel<BM_THREADS_WIDE, true, false>, BM_THREADS_WIDE, 0) : launch(bmGenerateKernel<BM_THREADS_WIDE, false, false>, BM_THREADS_WIDE, 1);
			else rc = tma ? launch(bmGenerateKernel<BM_THREADS, true, false>, BM_THREADS, 2) : launch(bmGenerateKernel<BM_THREADS, false, false>, BM_THREADS, 3);
		}
	}
	if (rc == FMB_OK) {
		for (uint64_t i = 0; i < TF; i++) out[i] = newView(slab, (double*)slab->base + i * paths, paths);
	} else if (slab) {
		poolFree(slab->base, slab->bytes);
		delete slab;
	}
	if (dsq) poolFree(dsq, TF * sizeof(double));
	poolFree(heads, (size_t)B * MT_N * sizeof(uint32_t));
	return rc;
}

int fmb_bm_generate(int32_t seed, int T, int F, uint64_t paths, uint64_t path_offset, const double* sqrt_dt, fmb_handle* out) {
	return generateIncrements((int64_t)seed, T, F, paths, path_offset, sqrt_dt, false, out);
}

int fmb_uniforms_generate(int64_t seed, int T, int F, uint64_t paths, uint64_t path_offset, fmb_handle* out) {
	return generateIncrements(seed, T, F, paths, path_offset, nullptr, true, out);
}

} // extern "C"
