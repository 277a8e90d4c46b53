/* Bit-sink plug-in point of the host encoder.
 *
 * ABI-identical to the reference's src/encoder_interface.h:7-13 so that existing back ends
 * (a cost counter, a range coder) keep working unchanged against this host:
 *   encode_bit(enc, bit, prob)           one modelled bit; prob = P(bit == 0) * 2048
 *   encode_direct_bits(enc, bits, n)     n equiprobable bits, most significant first
 * The struct is owned by the caller; private_data belongs to whoever filled the pointers.
 * The CUDA library does not call through this per bit: it implements the cost back end in
 * bulk (mg_score_slabs) and the range-coder back end on the device (mg_encode_slab).
 */
#ifndef MEGALANIA_ENCODER_INTERFACE_H
#define MEGALANIA_ENCODER_INTERFACE_H

#include <stdbool.h>
#include <stdint.h>

#ifndef Prob
#define Prob uint16_t
#endif

typedef struct EncoderInterface_struct EncoderInterface;

struct EncoderInterface_struct {
	void (*encode_bit)(EncoderInterface* enc, bool bit, Prob prob);
	void (*encode_direct_bits)(EncoderInterface* enc, unsigned bits, unsigned num_bits);
	void* private_data;
};

#endif
