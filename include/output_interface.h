/* Byte-sink plug-in point of the host encoder.
 *
 * ABI-identical to the reference's src/output_interface.h:8-13: write() receives a run of
 * finished output bytes and returns false on failure.  mg_encode_slab() hands the 13-byte
 * .lzma header and then the range-coded payload to this interface, so a FILE*, a memory
 * buffer or a socket sink written for the reference is a drop-in here.
 */
#ifndef MEGALANIA_OUTPUT_INTERFACE_H
#define MEGALANIA_OUTPUT_INTERFACE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct OutputInterface_struct OutputInterface;

struct OutputInterface_struct {
	bool (*write)(OutputInterface* out, const void* data, size_t data_size);
	void* private_data;
};

#endif
