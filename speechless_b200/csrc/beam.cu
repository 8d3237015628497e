// beam.cu — CTC prefix beam search, one CTA per utterance.
//
// Replaces tf.nn.ctc_beam_search_decoder as the reference calls it (net.py:444-451) with the STOCK
// scorer; the KenLM scorer lives in a patched TensorFlow fork (net.py:420-422) that is not in the
// reference tree.  Semantics restated from TensorFlow's ctc_beam_search.h (see
// oracle/beam_search_oracle.py for the CPU restatement and its pinning):
//   * the beam holds up to W label prefixes, each with log P(prefix, last frame blank) `pb`,
//     log P(prefix, last frame = its last label) `pl` and their log-sum `tot`;
//   * per frame every active prefix i is continued
//       pl' = logaddexp(pl, parent active ? (label_i == label_parent ? pb_parent : tot_parent) : 0) + lp[label_i]
//       pb' = tot + lp[blank],
//     and spawns the children (i, c) that are not already in the beam with
//       pl = lp[c] + (c == label_i ? pb_i : tot_i),  pb = 0;
//     the W best of {continued prefixes} U {children} survive.  TF's Step additionally wipes a prefix
//     that is displaced before the child loop reaches its parent, which silently drops that prefix's
//     own children for the frame (`tf_deactivation` in oracle/beam_search_oracle.py; it cannot occur for
//     beam_width = 1).  The default mode reproduces this order-dependent behaviour literally
//     (tf_exact_child_loop); SL_BEAM_ORDER_INDEPENDENT=1 selects the parallel "W best of the union" rule.
//   * merge_repeated only post-processes the output (a label equal to its successor's is dropped),
//     which is what makes "A A _ A A" decode to [0] / [0, 0] (reference test_ctc_decoders.py:38-39).
//
// Data layout: the beam lives in shared memory (two generations of W slots); the W*V candidate scores
// of a frame stay in registers (4 per thread) as order-preserving integers, the score of the W-th best
// is found by a bit-wise binary search (33 block-wide counts; a full bitonic sort of 4096 64-bit keys
// in shared memory was measured at 58 us per frame, this is ~10x less), the <= W survivors are ranked
// by counting; prefixes are nodes (parent, label) of a tree in global
// memory (a fresh node id = 1 + frame*W + slot, so no allocator is needed), walked back once at the
// end.  A prefix keeps its node when it drops out of the beam and comes back later — TF keeps the
// BeamEntry object in its tree, and its surviving children still point at it — so (parent node, label)
// -> node is kept in an open-addressing hash table in global memory, and parent slots are re-derived
// from node ids every frame.
//
// Word language model inside the search (sl_ctc_beam_search_decode_lm; the reference's KenLM branch,
// net.py:444-451): the four scorer hooks of TF's decoder are implemented on the device.  Every beam slot
// carries the scorer state of its prefix (trie node of the unfinished word, the last order - 1 finished
// words, the weighted LM total, the score including the look-ahead of the unfinished word, and the delta to
// the state it was expanded from); a state is a function of the prefix alone, so a prefix that re-enters the
// beam gets it re-derived from its parent and nothing is stored per tree node.  Per frame the deltas of all
// beam_width x V possible children are computed in parallel (letters: two reads of the vocabulary trie; the
// space label: a back-off walk through the n-gram hash table), then the selection runs as before on
// lp + previous mass + delta.  Restated from oracle/beam_search_oracle.py (WordLanguageModelScorer; parity
// with the patched TensorFlow fork is unpinned, see there).
#include <cstdlib>

#include "../../include/speechless_b200.h"
#include "common.cuh"

namespace sl {

namespace {

constexpr int BS_THREADS = 1024;  // (256 threads x 16 candidates each measured slower: 8.2 vs 7.2 ms at width 100)
constexpr int BS_MAX_W = 128;
constexpr int BS_MAX_CAND = 4096;  // beam_width * V candidates per frame
constexpr int BS_VP = 64;

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ float logaddexp(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(fminf(a, b) - m));
}

constexpr int LM_CTX = 4;  // finished words a state remembers: n-gram order <= 5
struct LmState {
  int trie;          // node of the unfinished word in the vocabulary trie: 0 = no letters yet, -1 = not a prefix of any word
  int ctx[LM_CTX];   // ids of the last finished words, newest last, -1 = none
  float total;       // weighted LM total of the finished words (incl. word bonuses)
  float score;       // total + look-ahead of the unfinished word
  float delta;       // score - score of the state this one was expanded from (TF: GetStateExpansionScore - previous)
};

// ---- n-gram table: open addressing, 8 ints per slot {n, id0..id4 (-1 padded), log10 p, log10 back-off};
//      slot = FNV-1a over the n ids, linear probing (speechless_b200/language_model.py builds it)
__device__ __forceinline__ bool lm_find(const SlWordLm& lm, const int* ids, int n, float* logp, float* backoff) {
  uint32_t h = 2166136261u;
  for (int k = 0; k < n; ++k) h = (h ^ static_cast<uint32_t>(ids[k])) * 16777619u;
  for (uint32_t slot = h & lm.ngram_mask;; slot = (slot + 1) & lm.ngram_mask) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(lm.ngrams) + 2 * slot);
    if (a.x == 0) return false;
    if (a.x != n) continue;
    const int4 b = __ldg(reinterpret_cast<const int4*>(lm.ngrams) + 2 * slot + 1);
    const int key[5] = {a.y, a.z, a.w, b.x, b.y};
    bool same = true;
    for (int k = 0; k < n; ++k) same = same && key[k] == ids[k];
    if (!same) continue;
    *logp = __int_as_float(b.z);
    *backoff = __int_as_float(b.w);
    return true;
  }
}
// log10 P(word | ctx) with ARPA back-off semantics
__device__ float lm_word_log10(const SlWordLm& lm, const int* ctx, int word) {
  if (word == lm.unk_id && !lm.has_unk) return lm.unknown_log10;
  int ids[LM_CTX + 1];
  int m = 0;
  for (int k = 0; k < LM_CTX; ++k)
    if (ctx[k] >= 0 && LM_CTX - k <= lm.order - 1) ids[m++] = ctx[k];
  ids[m] = word;
  float penalty = 0.f;
  for (int s0 = 0;; ++s0) {
    float logp, backoff;
    if (lm_find(lm, ids + s0, m - s0 + 1, &logp, &backoff)) return penalty + logp;
    if (s0 == m) return penalty + lm.unknown_log10;
    if (lm_find(lm, ids + s0, m - s0, &logp, &backoff)) penalty += backoff;
  }
}
// a finished word: what it adds to the total, and the new context
__device__ float lm_finish_word(const SlWordLm& lm, int trie, int* ctx) {
  const int word = trie > 0 ? __ldg(lm.trie_word + trie) : -1;
  const bool known = word >= 0;
  const int id = known ? word : lm.unk_id;
  const float gained = lm.weight * lm_word_log10(lm, ctx, id) + lm.word_count_weight +
                       (known ? lm.valid_word_count_weight : 0.f);
  for (int k = 0; k + 1 < LM_CTX; ++k) ctx[k] = ctx[k + 1];
  ctx[LM_CTX - 1] = id;
  return gained;
}
// TF ExpandState: the state of prefix + label
__device__ LmState lm_expand(const SlWordLm& lm, const LmState& from, int label, int V) {
  LmState to = from;
  if (label != lm.space_label) {
    to.trie = from.trie >= 0 ? __ldg(lm.trie_children + static_cast<size_t>(from.trie) * V + label) : -1;
    to.score = from.total + lm.weight * (to.trie >= 0 ? __ldg(lm.trie_min_unigram + to.trie) : lm.unknown_log10);
  } else {
    to.total = from.total + lm_finish_word(lm, from.trie, to.ctx);
    to.score = to.total;
    to.trie = 0;
  }
  to.delta = to.score - from.score;
  return to;
}
// TF ExpandStateEnd + GetStateEndExpansionScore: the unfinished word and the end of the sentence
__device__ float lm_end_delta(const SlWordLm& lm, const LmState& s) {
  int ctx[LM_CTX];
  for (int k = 0; k < LM_CTX; ++k) ctx[k] = s.ctx[k];
  float total = s.total;
  if (s.trie != 0) total += lm_finish_word(lm, s.trie, ctx);
  total += lm.weight * lm_word_log10(lm, ctx, lm.eos_id);
  return total - s.score;
}

struct BeamGen {  // one generation of the beam (slots sorted best first)
  int node[BS_MAX_W];
  int label[BS_MAX_W];
  int pslot[BS_MAX_W];  // slot of the parent prefix in the same generation, -1 = not in the beam
  float pb[BS_MAX_W];
  float pl[BS_MAX_W];
  float tot[BS_MAX_W];
  LmState st[BS_MAX_W];  // scorer state of the slot's prefix (language-model decode only)
};

struct BeamSmem {
  BeamGen gen[2];
  float pbn[BS_MAX_W], pln[BS_MAX_W], totn[BS_MAX_W];  // continued prefixes of the current frame
  unsigned long long child_active[BS_MAX_W];           // bit c: child (slot, c) is already in the beam
  unsigned long long selected[BS_MAX_W];               // keys of the survivors of a frame, unordered
  int parent_node[BS_MAX_W];                           // node id of the parent of next-generation slot r
  float lp[BS_VP];
  int partial[2][32];  // per-warp partial counts of the selection (BS_THREADS / 32 warps)
  int n_selected;
  int n_active;
  // TF-exact mode (sequential child loop of TensorFlow's CTCBeamSearchDecoder::Step, run by warp 0)
  int leaf_idx[BS_MAX_W];                // `leaves_` of the frame being built: candidate index slot * V + symbol
                                         // (symbol == blank: prefix `slot` continued); the scores sit in registers
  unsigned char child_slot[BS_MAX_W][BS_VP];  // slot of the beam entry that is child (slot, symbol), valid where child_active
  unsigned char displaced[BS_MAX_W];     // beam entry popped from `leaves_` during this frame's child loop
  unsigned char wiped[BS_MAX_W];         // ... and then "deactivated" through its parent: it proposes no children
  float cand_delta[BS_MAX_W][BS_VP];     // language model: delta of the state of child (slot, symbol)
};

constexpr int BS_CPT = BS_MAX_CAND / BS_THREADS;  // candidates per thread, kept in registers
constexpr uint32_t ORD_NEG_INF = 0x007fffffu;     // float_to_ordered(-inf); 0 marks "no candidate"

// Block-wide sum of a small per-thread count, without atomics: warp sums go to one of two slot rows
// (`it` is a per-thread copy of a block-uniform call counter; call n uses row n & 1, whose readers of
// call n - 2 are all past the barrier of call n - 1), one barrier, every warp adds up the 32 slots.
__device__ __forceinline__ int block_sum(BeamSmem& sm, int c, int& it) {
  int* part = sm.partial[it & 1];
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  ++it;
  const int lane = threadIdx.x & 31;
  return __reduce_add_sync(0xffffffffu, lane < BS_THREADS / 32 ? part[lane] : 0);
}

// TensorFlow's CTCBeamSearchDecoder::Step, child loop included, restated literally (ctc_beam_search.h; CPU
// restatement with the derivation: oracle/beam_search_oracle.py, tf_deactivation = True).  `leaves_` starts as the
// continued prefixes of the beam.  Then, for every beam entry b in order of its previous total (= slot order)
// whose previous total still beats the bottom of `leaves_` (and has not been wiped), and every label c in
// order: the child (b, c) is skipped if it is in `leaves_`; otherwise it is scored afresh
// (lp[c] + previous mass of b) and, if it beats the bottom, replaces it.  The order-dependent part: a beam
// entry X that has been replaced ("displaced") before the loop reaches its parent is re-scored as if new when
// the parent gets to it, cannot beat the bottom that displaced it, and has its previous probabilities wiped —
// so X proposes no children in this frame.  Inherently sequential (each insertion moves the bottom), but
// cheap: one warp, `leaves_` in shared memory, per beam entry one ballot of the labels whose score beats the
// bottom (it only rises) and a visit of those few.  Returns the number of leaves; their keys go to
// sm.selected in the format of the parallel selection.  All threads call it (barriers inside).
__device__ __noinline__ int tf_exact_child_loop(BeamSmem& sm, const BeamGen& g, int n, int V, int blank, int W,
                                                bool with_lm) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < n) sm.leaf_idx[tid] = tid * V + blank;
  __syncthreads();
  if (tid < 32) {
    // The leaves: slot k = lane + 32 r lives in register key[r] of its lane as the order-preserving integer
    // image of the score (an empty slot holds the largest key, so it is never the bottom); the candidate
    // index of a slot sits in shared memory.  Bottom = one min per lane, one redux, one ballot.
    constexpr int R = BS_MAX_W / 32;
    uint32_t key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) key[r] = lane + 32 * r < n ? float_to_ordered(sm.totn[lane + 32 * r]) : 0xffffffffu;
    int count = n;  // (continued prefixes of probability zero stay, as in TF, and are dropped at the end)
    uint32_t bottom = 0u;
    int bottom_pos = -1;
    auto find_bottom = [&]() {
      uint32_t my_min = key[0];
      int my_pos = lane;
#pragma unroll
      for (int r = 1; r < R; ++r)
        if (key[r] <= my_min && lane + 32 * r < count) {
          my_min = key[r];
          my_pos = lane + 32 * r;
        }
      bottom = __reduce_min_sync(0xffffffffu, my_min);
      const unsigned who = __ballot_sync(0xffffffffu, my_min == bottom);
      bottom_pos = __shfl_sync(0xffffffffu, my_pos, 31 - __clz(who));
    };
    find_bottom();
    for (int i = 0; i < n; ++i) {
      const float old_total = g.tot[i];
      // is_candidate(b.old): not wiped, finite, and (beam not full or better than the bottom).  The entries come
      // in order of decreasing previous total and the bottom only rises: once a full beam's bottom has caught up
      // with an entry's previous total, no later entry can propose anything either.
      if (count == W && !(float_to_ordered(old_total) > bottom)) break;
      if (sm.wiped[i] || !(old_total > -INFINITY)) continue;
      const int lab_i = g.label[i];
      const unsigned long long in_beam = sm.child_active[i];
      // per lane: labels lane and lane + 32
      uint32_t sc[2];
      unsigned beam_children[2], todo[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        sc[h] = 0u;
        if (c < V && c != blank)
          sc[h] = float_to_ordered(sm.lp[c] + (c == lab_i ? g.pb[i] : g.tot[i]) + (with_lm ? sm.cand_delta[i][c] : 0.f));
        beam_children[h] = static_cast<unsigned>(in_beam >> (32 * h));
        // worth a visit: a candidate that beats the bottom (which only rises), or a beam entry — it may have
        // been displaced by now
        todo[h] = __ballot_sync(0xffffffffu, sc[h] > ORD_NEG_INF && (count < W || sc[h] > bottom)) | beam_children[h];
      }
      for (int h = 0; h < 2; ++h) {
        while (todo[h]) {
          const int l = __ffs(todo[h]) - 1;
          todo[h] &= todo[h] - 1;
          const int c = l + 32 * h;
          const uint32_t score = __shfl_sync(0xffffffffu, sc[h], l);
          const bool beam_child = (beam_children[h] >> l) & 1u;
          const int j = beam_child ? sm.child_slot[i][c] : -1;
          if (beam_child && !sm.displaced[j]) continue;  // active: already among the leaves
          const bool candidate = score > ORD_NEG_INF && (count < W || score > bottom);
          if (!candidate) {
            if (beam_child && lane == 0) sm.wiped[j] = 1;  // "deactivate child"
            __syncwarp();
            continue;
          }
          int pos = count;
          if (count == W) {
            pos = bottom_pos;
            const int gone = sm.leaf_idx[pos];
            const int gi = gone / V;
            if (gone - gi * V == blank && lane == 0) sm.displaced[gi] = 1;  // a continued beam entry leaves
          } else {
            ++count;
          }
          __syncwarp();
          if (lane == (pos & 31)) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r == (pos >> 5)) key[r] = score;
            sm.leaf_idx[pos] = i * V + c;
          }
          if (beam_child && lane == 0) sm.displaced[j] = 0;  // it is back among the leaves (as a fresh entry)
          __syncwarp();
          find_bottom();
          if (count == W) {
            // the bottom has risen: drop the candidates of this entry that can no longer beat it
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
              todo[hh] &= __ballot_sync(0xffffffffu, sc[hh] > bottom) | beam_children[hh];
          }
        }
      }
    }
    // keys of the surviving finite leaves, in the format of the parallel selection
    int kept = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int k = lane + 32 * r;
      const bool ok = k < count && key[r] > ORD_NEG_INF;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int at = kept + __popc(m & ((1u << lane) - 1));
        sm.selected[at] = (static_cast<unsigned long long>(key[r]) << 32) |
                          static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(sm.leaf_idx[k]));
      }
      kept += __popc(m);
    }
    if (lane == 0) sm.n_selected = kept;
  }
  __syncthreads();
  return sm.n_selected;
}

template <bool LM>
__global__ void __launch_bounds__(BS_THREADS)
    ctc_beam_search_kernel(const float* __restrict__ scores, const int32_t* __restrict__ input_len,
                           int32_t* __restrict__ out, int32_t* __restrict__ out_len, float* __restrict__ out_logp,
                           int2* __restrict__ nodes_all, unsigned long long* __restrict__ hash_all, int hash_cap,
                           int T, int V, int blank, int W, int top_paths, int merge_repeated, int inputs_are_probs,
                           int tf_exact, const SlWordLm lm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BeamSmem& sm = *reinterpret_cast<BeamSmem*>(smem_raw);
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int P = min(input_len[b], T);
  int2* nodes = nodes_all + static_cast<size_t>(b) * (static_cast<size_t>(T) * W + 1);  // (parent node, label)
  unsigned long long* hash = hash_all + static_cast<size_t>(b) * hash_cap;  // zeroed by the launcher
  const float* sc = scores + static_cast<size_t>(b) * T * V;

  if (tid == 0) {
    nodes[0] = make_int2(-1, -1);  // root: the empty prefix, P = 1 with "last frame blank"
    BeamGen& g = sm.gen[0];
    g.node[0] = 0;
    g.label[0] = -1;
    g.pslot[0] = -1;
    g.pb[0] = 0.f;
    g.pl[0] = -INFINITY;
    g.tot[0] = 0.f;
    if constexpr (LM) {  // TF InitializeState: nothing typed, history = <s>
      g.st[0].trie = 0;
      for (int k = 0; k < LM_CTX; ++k) g.st[0].ctx[k] = k == LM_CTX - 1 ? lm.bos_id : -1;
      g.st[0].total = g.st[0].score = g.st[0].delta = 0.f;
    }
    sm.n_active = 1;
  }
  // the frame's scores are fetched one frame ahead (warp 0, two symbols per lane)
  auto fetch = [&](int t, float& x0, float& x1) {
    x0 = x1 = -INFINITY;
    if (tid < 32 && t < P) {
      const float* row = sc + static_cast<size_t>(t) * V;
      if (tid < V) x0 = row[tid];
      if (tid + 32 < V) x1 = row[tid + 32];
    }
  };
  float x0, x1;
  fetch(0, x0, x1);
  __syncthreads();

  int cur = 0, it = 0;
  const int n_cand = W * V;
  for (int t = 0; t < P; ++t) {
    BeamGen& g = sm.gen[cur];
    BeamGen& nx = sm.gen[cur ^ 1];
    const int n = sm.n_active;
    // 1. log-softmax of the frame (TF normalises every frame; the reference feeds log(p + 1e-8))
    if (tid < 32) {
      float y0 = x0, y1 = x1;
      if (inputs_are_probs) {
        y0 = tid < V ? logf(x0 + 1e-8f) : -INFINITY;
        y1 = tid + 32 < V ? logf(x1 + 1e-8f) : -INFINITY;
      }
      float m = fmaxf(y0, y1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = (tid < V ? expf(y0 - m) : 0.f) + (tid + 32 < V ? expf(y1 - m) : 0.f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float norm = m + logf(s);
      sm.lp[tid] = y0 - norm;
      sm.lp[tid + 32] = y1 - norm;
    }
    fetch(t + 1, x0, x1);
    if (tid < W) {
      sm.child_active[tid] = 0ull;
      sm.displaced[tid] = 0;
      sm.wiped[tid] = 0;
    }
    if (tid == 0) sm.n_selected = 0;
    __syncthreads();
    // 2. continue the active prefixes; mark which children are already in the beam
    if (tid < n) {
      const int lab = g.label[tid];
      float pl = -INFINITY;
      if (lab >= 0) {
        const int ps = g.pslot[tid];
        float previous = ps >= 0 ? (lab == g.label[ps] ? g.pb[ps] : g.tot[ps]) : -INFINITY;
        if constexpr (LM) previous += g.st[tid].delta;  // TF GetStateExpansionScore(b->state, previous)
        pl = logaddexp(g.pl[tid], previous) + sm.lp[lab];
        if (ps >= 0) {
          atomicOr(&sm.child_active[ps], 1ull << lab);
          sm.child_slot[ps][lab] = static_cast<unsigned char>(tid);
        }
      }
      const float pb = g.tot[tid] + sm.lp[blank];
      sm.pbn[tid] = pb;
      sm.pln[tid] = pl;
      sm.totn[tid] = logaddexp(pb, pl);
    }
    if constexpr (LM) {
      // the scorer deltas of every possible child (TF calls ExpandState inside its child loop; a state is a
      // function of the prefix alone, so computing them up front changes nothing)
      for (int idx = tid; idx < n * V; idx += BS_THREADS) {
        const int i = idx / V, c = idx - i * V;
        if (c != blank) sm.cand_delta[i][c] = lm_expand(lm, g.st[i], c, V).delta;
      }
    }
    __syncthreads();
    int n_next;
    if (tf_exact) {
      n_next = tf_exact_child_loop(sm, g, n, V, blank, W, LM);  // fills sm.selected[0 .. n_next)
    } else {
      // 3. candidates (slot i, symbol c), BS_CPT per thread in registers; c == blank stands for
      //    "prefix i continued".  Order-preserving integer image of the score; 0 = no candidate.
      uint32_t kv[BS_CPT];
#pragma unroll
      for (int q = 0; q < BS_CPT; ++q) {
        const int idx = tid + q * BS_THREADS;
        kv[q] = 0u;
        if (idx < n_cand) {
          const int i = idx / V, c = idx - i * V;
          if (i < n) {
            float key = -INFINITY;
            if (c == blank)
              key = sm.totn[i];
            else if (!((sm.child_active[i] >> c) & 1ull))
              key = sm.lp[c] + (c == g.label[i] ? g.pb[i] : g.tot[i]) + (LM ? sm.cand_delta[i][c] : 0.f);
            kv[q] = float_to_ordered(key);
          }
        }
      }
      // 4. the W best: binary search, bit by bit, for the score of the W-th best candidate (one block-wide
      //    count per bit); ties at that score go to the lower candidate index
      auto count_if = [&](auto pred) {
        int c = 0;
#pragma unroll
        for (int q = 0; q < BS_CPT; ++q) c += pred(kv[q], tid + q * BS_THREADS) ? 1 : 0;
        return block_sum(sm, c, it);
      };
      const int n_finite = count_if([](uint32_t k, int) { return k > ORD_NEG_INF; });
      n_next = min(W, n_finite);
      uint32_t thr = 0u;
      bool exact = false;  // some prefix of the threshold already separates exactly n_next candidates
      for (int bit = 31; bit >= 0 && !exact; --bit) {
        const uint32_t cand = thr | (1u << bit);
        const int c = count_if([cand](uint32_t k, int) { return k >= cand; });
        if (c >= n_next) thr = cand;
        exact = c == n_next;
      }
      int idx_limit = BS_MAX_CAND;  // ties with candidate index <= idx_limit are taken
      int n_greater = 0, n_ties = 0;
      if (exact) {
        thr -= 1u;        // "k >= thr" as "k > thr - 1" (thr > 0: it is at least the image of a finite score)
        idx_limit = -1;   // ...and nothing that merely equals thr - 1
      } else {
        n_greater = count_if([thr](uint32_t k, int) { return k > thr; });
        n_ties = count_if([thr](uint32_t k, int) { return k == thr; });
      }
      if (!exact && n_greater + n_ties > n_next) {
        const int need = n_next - n_greater;
        int lo = 0;
        for (int bit = 11; bit >= 0; --bit) {
          const int cand = lo | (1 << bit);
          if (count_if([thr, cand](uint32_t k, int idx) { return k == thr && idx < cand; }) < need) lo = cand;
        }
        idx_limit = lo;
      }
#pragma unroll
      for (int q = 0; q < BS_CPT; ++q) {
        const int idx = tid + q * BS_THREADS;
        if (kv[q] > thr || (kv[q] == thr && kv[q] > ORD_NEG_INF && idx <= idx_limit)) {
          const int pos = atomicAdd(&sm.n_selected, 1);
          sm.selected[pos] = (static_cast<unsigned long long>(kv[q]) << 32) |
                             static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(idx));
        }
      }
      __syncthreads();
    }
    // 5. survivor r takes the slot given by its rank (keys are unique: score, then lower index first)
    if (tid < n_next) {
      const unsigned long long key = sm.selected[tid];
      int slot = 0;
      for (int s2 = 0; s2 < n_next; ++s2) slot += sm.selected[s2] > key ? 1 : 0;
      const float f = ordered_to_float(static_cast<uint32_t>(key >> 32));
      const int idx = static_cast<int>(0xffffffffu - static_cast<uint32_t>(key & 0xffffffffull));
      const int i = idx / V, c = idx - i * V;
      if (c == blank) {
        const int node = g.node[i];
        nx.node[slot] = node;
        nx.label[slot] = g.label[i];
        nx.pb[slot] = sm.pbn[i];
        nx.pl[slot] = sm.pln[i];
        nx.tot[slot] = sm.totn[i];
        if constexpr (LM) nx.st[slot] = g.st[i];
        sm.parent_node[slot] = nodes[node].x;
      } else {
        // the prefix (node of i) + c: reuse its node if it has been in the beam before
        const int pn = g.node[i];
        const unsigned long long tag = (static_cast<unsigned long long>(pn) * BS_VP + c + 1ull) << 32;
        int id = 1 + t * W + slot;
        unsigned h = (static_cast<unsigned>(pn) * 2654435761u + static_cast<unsigned>(c) * 40503u) & (hash_cap - 1);
        for (;;) {
          const unsigned long long seen = atomicCAS(&hash[h], 0ull, tag | static_cast<unsigned>(id));
          if (seen == 0ull) {
            nodes[id] = make_int2(pn, c);
            break;
          }
          if ((seen >> 32) == (tag >> 32)) {
            id = static_cast<int>(seen & 0xffffffffull);
            break;
          }
          h = (h + 1) & (hash_cap - 1);
        }
        nx.node[slot] = id;
        nx.label[slot] = c;
        nx.pb[slot] = -INFINITY;
        nx.pl[slot] = f;
        nx.tot[slot] = f;
        if constexpr (LM) nx.st[slot] = lm_expand(lm, g.st[i], c, V);
        sm.parent_node[slot] = pn;
      }
    }
    __syncthreads();
    if (tid < n_next) {
      const int pn = sm.parent_node[tid];
      int ps = -1;
      for (int s2 = 0; s2 < n_next; ++s2)
        if (nx.node[s2] == pn) ps = s2;
      nx.pslot[tid] = ps;
    }
    if (tid == 0) sm.n_active = n_next;
    __syncthreads();
    cur ^= 1;
  }

  // output: walk the best prefixes back to the root (TF LabelSeq: with merge_repeated a label equal
  // to the label of the node visited just before it, i.e. its successor, is dropped)
  const BeamGen& g = sm.gen[cur];
  const int n = sm.n_active;
  if constexpr (LM) {
    // TF TopPaths: every leaf's state is closed (unfinished word, end of sentence) before the leaves are ranked
    if (tid < n) sm.totn[tid] = g.tot[tid] + lm_end_delta(lm, g.st[tid]);
    __syncthreads();
    if (tid < n) {
      int rank = 0;
      for (int s2 = 0; s2 < n; ++s2)
        rank += (sm.totn[s2] > sm.totn[tid] || (sm.totn[s2] == sm.totn[tid] && s2 < tid)) ? 1 : 0;
      sm.leaf_idx[rank] = tid;
    }
    __syncthreads();
  }
  if (tid < top_paths) {
    int32_t* o = out + (static_cast<size_t>(b) * top_paths + tid) * T;
    int len = 0;
    const int src = (LM && tid < n) ? sm.leaf_idx[tid] : tid;  // slot of the tid-th best finished hypothesis
    if (tid < n) {
      int prev = -1;
      for (int node = g.node[src]; node > 0; node = nodes[node].x) {
        const int lab = nodes[node].y;
        if (!merge_repeated || lab != prev) ++len;
        prev = lab;
      }
      int pos = len;
      prev = -1;
      for (int node = g.node[src]; node > 0; node = nodes[node].x) {
        const int lab = nodes[node].y;
        if (!merge_repeated || lab != prev) o[--pos] = lab;
        prev = lab;
      }
    }
    for (int i = len; i < T; ++i) o[i] = -1;
    out_len[b * top_paths + tid] = tid < n ? len : 0;
    out_logp[b * top_paths + tid] = tid < n ? (LM ? sm.totn[src] : g.tot[src]) : -INFINITY;
  }
}

}  // namespace

static int beam_hash_capacity(int T, int beam_width) {
  size_t cap = 64;
  while (cap < 2 * (static_cast<size_t>(T) * beam_width + 1)) cap <<= 1;
  return static_cast<int>(cap);
}
static size_t beam_nodes_bytes(int B, int T, int beam_width) {
  const size_t n = static_cast<size_t>(B) * (static_cast<size_t>(T) * beam_width + 1) * sizeof(int2);
  return (n + 255) & ~static_cast<size_t>(255);
}
size_t beam_search_workspace_bytes(int B, int T, int beam_width) {
  return beam_nodes_bytes(B, T, beam_width) +
         static_cast<size_t>(B) * beam_hash_capacity(T, beam_width) * sizeof(unsigned long long);
}

int beam_search_launch(const float* scores, const int32_t* input_len, int32_t* out, int32_t* out_len, float* out_logp,
                       int B, int T, int V, int blank, int beam_width, int top_paths, int merge_repeated,
                       int inputs_are_probs, const SlWordLm* lm, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  SL_REQUIRE(V >= 2 && V <= BS_VP, "beam search supports 2..64 symbols (incl. blank)");
  SL_REQUIRE(blank >= 0 && blank < V, "blank out of range");
  SL_REQUIRE(beam_width >= 1 && beam_width <= BS_MAX_W, "beam_width must be 1..128");
  SL_REQUIRE(beam_width * V <= BS_MAX_CAND, "beam_width * symbols must not exceed 4096");
  SL_REQUIRE(top_paths >= 1 && top_paths <= beam_width, "top_paths must be 1..beam_width");
  SL_REQUIRE(workspace_bytes >= beam_search_workspace_bytes(B, T, beam_width), "beam search workspace too small");
  SL_REQUIRE(static_cast<size_t>(T) * beam_width < (1u << 24), "too many frames x beam entries");
  const size_t smem = sizeof(BeamSmem);
  if (lm != nullptr) {
    SL_REQUIRE(lm->trie_children && lm->trie_word && lm->trie_min_unigram && lm->ngrams, "language model: null table");
    SL_REQUIRE(lm->order >= 1 && lm->order <= LM_CTX + 1, "language model: n-gram order must be 1..5");
    SL_REQUIRE(lm->n_labels == V, "language model: the trie must have one column per symbol");
    SL_REQUIRE(lm->space_label >= 0 && lm->space_label < V && lm->space_label != blank, "language model: bad space label");
    SL_REQUIRE((lm->ngram_mask & (lm->ngram_mask + 1)) == 0, "language model: the table size must be a power of two");
    SL_CUDA(cudaFuncSetAttribute(ctc_beam_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  } else {
    SL_CUDA(cudaFuncSetAttribute(ctc_beam_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  }
  const int hash_cap = beam_hash_capacity(T, beam_width);
  unsigned long long* hash = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) +
                                                                   beam_nodes_bytes(B, T, beam_width));
  SL_CUDA(cudaMemsetAsync(hash, 0, static_cast<size_t>(B) * hash_cap * sizeof(unsigned long long), stream));
  const char* oi = std::getenv("SL_BEAM_ORDER_INDEPENDENT");  // 1: parallel selection rule (A/B, see the header comment)
  // (with a language model always TF's loop: a word bonus can lift a child above its parent, and then "the W
  // best of the union" also takes children of parents that TF's is_candidate(parent) test never expands)
  const int tf_exact = (oi && std::atoi(oi) != 0 && lm == nullptr) ? 0 : 1;
  if (lm != nullptr)
    ctc_beam_search_kernel<true><<<B, BS_THREADS, smem, stream>>>(scores, input_len, out, out_len, out_logp,
                                                                  reinterpret_cast<int2*>(workspace), hash, hash_cap,
                                                                  T, V, blank, beam_width, top_paths, merge_repeated,
                                                                  inputs_are_probs, tf_exact, *lm);
  else
    ctc_beam_search_kernel<false><<<B, BS_THREADS, smem, stream>>>(scores, input_len, out, out_len, out_logp,
                                                                   reinterpret_cast<int2*>(workspace), hash, hash_cap,
                                                                   T, V, blank, beam_width, top_paths, merge_repeated,
                                                                   inputs_are_probs, tf_exact, SlWordLm{});
  SL_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sl
