/* TEST INFRASTRUCTURE (oracle/) - never linked into the product.
 *
 * Restatement of the convolutional decoder the reference's aero-decode gets from libcorrect
 * (quiet/libcorrect, pinned commit f5a28c74... in the reference README; NOT vendored under /root/reference), behind
 * the slice of its C API that decode/jconvolutionalcodec.cpp:10-16,30,64,97,164,233 calls. It exists so that the
 * UNMODIFIED decode sources link into oracle/_ref/libref_decode.so for the end-to-end check of SURVEY.md section 8d.
 *
 * PARITY UNPINNED: libcorrect's source is not available here, so this follows its published algorithm:
 *   - shift register with the newest bit in the LSB, output bit j = parity(register & poly[j]), output j is the j-th
 *     received symbol of a set;
 *   - hard metric = Hamming distance; soft metric = sum |soft - (bit ? 255 : 0)| (libcorrect's default "linear" metric);
 *   - the path starts in the all-zero state and is forced back to it by (order-1) zero inputs at the end of the
 *     buffer (libcorrect's warm-up and tail phases); one decoded bit per set, packed MSB first; the return value is
 *     the number of bytes written.
 * Difference: this decoder traces back once over the whole buffer, libcorrect traces back in windows (5*order minimum,
 * 15*order group). On the buffers jconvolutionalcodec.cpp hands over (a few hundred sets) the two agree unless the
 * survivor paths have not merged within 5*order steps, i.e. at very low SNR. The end-to-end test compares the
 * product's payloads with the CPU chain's THROUGH THE SAME decoder, so this does not weaken that comparison; it only
 * means absolute decode sensitivity is not claimed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include "shim_decode/correct.h"

struct correct_convolutional {
  unsigned rate, order;
  unsigned* table; /* [1 << order] outputs per full register value */
};

correct_convolutional* correct_convolutional_create(size_t inv_rate, size_t order, const correct_convolutional_polynomial_t* poly) {
  if (inv_rate < 2 || inv_rate > 8 || order < 2 || order > 15) return NULL;
  correct_convolutional* c = (correct_convolutional*)calloc(1, sizeof *c);
  c->rate = (unsigned)inv_rate;
  c->order = (unsigned)order;
  c->table = (unsigned*)malloc(sizeof(unsigned) << order);
  for (unsigned r = 0; r < (1u << order); r++) {
    unsigned out = 0;
    for (unsigned j = 0; j < c->rate; j++) out |= (unsigned)(__builtin_popcount(r & poly[j]) & 1) << j;
    c->table[r] = out;
  }
  return c;
}

void correct_convolutional_destroy(correct_convolutional* c) {
  if (!c) return;
  free(c->table);
  free(c);
}

size_t correct_convolutional_encode_len(correct_convolutional* c, size_t msg_len) { return c->rate * (8 * msg_len + c->order + 1); }

size_t correct_convolutional_encode(correct_convolutional* c, const uint8_t* msg, size_t msg_len, uint8_t* encoded) {
  const size_t nbits = correct_convolutional_encode_len(c, msg_len);
  memset(encoded, 0, (nbits + 7) / 8);
  unsigned reg = 0;
  const unsigned mask = (1u << c->order) - 1;
  size_t w = 0;
  for (size_t i = 0; i < 8 * msg_len + c->order + 1; i++) {
    const unsigned bit = i < 8 * msg_len ? (msg[i >> 3] >> (7 - (i & 7))) & 1u : 0u;
    reg = ((reg << 1) | bit) & mask;
    const unsigned out = c->table[reg];
    for (unsigned j = 0; j < c->rate; j++, w++)
      if ((out >> j) & 1u) encoded[w >> 3] |= (uint8_t)(0x80u >> (w & 7));
  }
  return nbits;
}

static ssize_t decode_any(correct_convolutional* c, const uint8_t* enc, size_t num_encoded_bits, uint8_t* msg, int soft) {
  const size_t sets = num_encoded_bits / c->rate;
  if (sets == 0) return -1;
  const unsigned S = 1u << (c->order - 1); /* state = the previous order-1 input bits */
  const unsigned INF = 0x3fffffffu;
  unsigned* cur = (unsigned*)malloc(sizeof(unsigned) * S);
  unsigned* nxt = (unsigned*)malloc(sizeof(unsigned) * S);
  uint8_t* prev_msb = (uint8_t*)malloc(sets * (size_t)S); /* survivor: the bit shifted out of the predecessor */
  for (unsigned s = 0; s < S; s++) cur[s] = INF;
  cur[0] = 0;
  unsigned dist[256]; /* branch metric per output word of this set (rate <= 8) */
  for (size_t t = 0; t < sets; t++) {
    for (unsigned o = 0; o < (1u << c->rate); o++) {
      unsigned d = 0;
      for (unsigned j = 0; j < c->rate; j++) {
        const unsigned want = (o >> j) & 1u;
        if (soft) {
          const int y = enc[t * c->rate + j];
          const int x = want ? 255 : 0;
          d += (unsigned)(y > x ? y - x : x - y);
        } else {
          const size_t b = t * c->rate + j;
          d += (((enc[b >> 3] >> (7 - (b & 7))) & 1u) != want);
        }
      }
      dist[o] = d;
    }
    const int forced_zero = t + (c->order - 1) >= sets; /* tail: only zero inputs */
    for (unsigned ns = 0; ns < S; ns++) {
      /* next state ns = (reg & (S-1)) with reg = (ps << 1 | bit); ps is ns >> 1 with either MSB */
      const unsigned bit = ns & 1u;
      unsigned best = INF;
      uint8_t bm = 0;
      if (!(forced_zero && bit)) {
        for (unsigned m = 0; m < 2; m++) {
          const unsigned ps = (ns >> 1) | (m << (c->order - 2));
          if (cur[ps] >= INF) continue;
          const unsigned reg = (ps << 1) | bit;
          const unsigned v = cur[ps] + dist[c->table[reg]];
          if (v < best) {
            best = v;
            bm = (uint8_t)m;
          }
        }
      }
      nxt[ns] = best;
      prev_msb[t * S + ns] = bm;
    }
    unsigned* sw = cur;
    cur = nxt;
    nxt = sw;
  }
  /* trace back from the zero state (buffers shorter than order-1 sets: from the best state) */
  unsigned s = 0;
  if (cur[0] >= INF) {
    unsigned best = INF;
    for (unsigned k = 0; k < S; k++)
      if (cur[k] < best) {
        best = cur[k];
        s = k;
      }
  }
  const size_t nbytes = (sets + 7) / 8;
  memset(msg, 0, nbytes);
  for (size_t t = sets; t-- > 0;) {
    if (s & 1u) msg[t >> 3] |= (uint8_t)(0x80u >> (t & 7));
    s = (s >> 1) | ((unsigned)prev_msb[t * S + s] << (c->order - 2));
  }
  free(cur);
  free(nxt);
  free(prev_msb);
  return (ssize_t)nbytes;
}

ssize_t correct_convolutional_decode(correct_convolutional* c, const uint8_t* encoded, size_t num_encoded_bits, uint8_t* msg) {
  return decode_any(c, encoded, num_encoded_bits, msg, 0);
}

ssize_t correct_convolutional_decode_soft(correct_convolutional* c, const correct_convolutional_soft_t* encoded, size_t num_encoded_bits, uint8_t* msg) {
  return decode_any(c, encoded, num_encoded_bits, msg, 1);
}
