/* Host-side I/O of the drop-in CLI: the input file as read-only memory, stdout as the
 * OutputInterface sink mg_encode_slab writes the .lzma stream to.  Plain host code: it plays the
 * part of the reference's file_output / memory_mapper pair behind the same plug-in interface. */
#ifndef MEGALANIA_HOST_IO_H
#define MEGALANIA_HOST_IO_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "output_interface.h"

typedef struct {
	FILE* stream;
	size_t bytes_written;
	int failed; /* sticky: a failed write fails every later one */
} StreamSink;

/* Fills `output` so that its write() appends to `stream`; `sink` is borrowed (caller-owned). */
void stream_sink_init(OutputInterface* output, StreamSink* sink, FILE* stream);

typedef struct {
	const uint8_t* data;
	size_t size;
} InputFile;

/* Maps `path` read-only.  0 on success; -1 with a line on stderr otherwise (a zero-length file
 * cannot be mapped and is reported the way the reference reports it). */
int input_file_open(InputFile* in, const char* path);
void input_file_close(InputFile* in);

#endif
