/* Read-only file mapping for the host CLI (the role of the reference's src/memory_mapper.{c,h}). */
#ifndef MEGALANIA_MEMORY_MAPPER_H
#define MEGALANIA_MEMORY_MAPPER_H
#include <stddef.h>
#include <stdint.h>

/* 0 on success, -1 (with a message on stderr) on failure */
int map_file(const char* filename, const uint8_t** data, size_t* data_size);
int unmap(const uint8_t* data, size_t data_size);

#endif
