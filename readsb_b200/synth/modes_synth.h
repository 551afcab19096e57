/*
 * modes_synth.h — deterministic synthetic 2.4 MSPS uc8 IQ captures with Mode-S replies injected.
 *
 * The reference ships no recorded capture, so the benchmark and parity inputs
 * are generated: DF17/DF11/... frames with valid CRC-24 parity, 12 MHz-tick start offsets (all five
 * sub-sample phases), per-frame amplitude and carrier phase, additive noise, overlap by complex
 * addition and clipping to the 8-bit range.  Pulse model as SURVEY.md section 8d: preamble pulses at
 * 0, 1.0, 3.5, 4.5 us (0.5 us wide), data from 8 us, PPM bit 1 = high then low; one sample =
 * 5 ticks of the 12 MHz clock, sample value = amplitude * fraction of the sample that is high.
 * Pure integer / un-contracted float arithmetic with an own PRNG: same bytes on every host.
 */
#ifndef MODES_SYNTH_H
#define MODES_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SYNTH_DF17   0x01u  /* extended squitter, PI = 0 overlay */
#define SYNTH_DF11   0x02u  /* all-call reply, II = 0 */
#define SYNTH_AP     0x04u  /* DF0/4/5/16/20/21 with address/parity */
#define SYNTH_DF18   0x08u
#define SYNTH_DF11_IID 0x10u /* DF11 with a non-zero interrogator id in the parity */
#define SYNTH_MODEAC 0x20u  /* Mode A/C reply: F1, 12 code pulses, F2 (+SPI sometimes), 0.45 us pulses every 1.45 us */

typedef struct synth_params {
    uint64_t seed;
    double   frames_per_sec;   /* injected replies per second of stream time */
    uint32_t df_mask;          /* SYNTH_* mix, frames cycle through the enabled kinds */
    uint32_t n_icao;           /* distinct aircraft addresses (derived from seed) */
    double   amp_min, amp_max; /* fraction of full scale */
    double   noise_sigma;      /* LSB per I and Q component */
    double   p_bit_error;      /* probability that a frame gets one flipped bit */
    double   p_two_bit_error;  /* probability that a frame gets two flipped bits */
} synth_params;

typedef struct synth_truth {
    int64_t  start_tick;       /* 12 MHz tick of the first preamble pulse */
    uint8_t  msg[14];          /* bytes as transmitted (after error injection) */
    uint8_t  nbits;            /* 56 / 112 */
    uint8_t  errors;           /* flipped bits */
} synth_truth;

/* Fill iq[2*nsamples].  truth (optional) receives up to truth_cap injected frames. Returns number
 * of frames injected. */
long synth_generate_uc8(const synth_params *p, uint64_t nsamples, uint8_t *iq,
                        synth_truth *truth, uint64_t truth_cap);

/* Mode-S CRC-24 parity (generator 0xFFF409) over the first nbits-24 bits, for building frames. */
uint32_t synth_crc24(const uint8_t *msg, int nbits_total);

#ifdef __cplusplus
}
#endif
#endif
