#include "memory_mapper.h"

#include <fcntl.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

int map_file(const char* filename, const uint8_t** data, size_t* data_size)
{
	int fd = open(filename, O_RDONLY);
	if (fd < 0) {
		perror(filename);
		return -1;
	}
	struct stat st;
	if (fstat(fd, &st) < 0) {
		perror("fstat");
		close(fd);
		return -1;
	}
	void* mem = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (mem == MAP_FAILED) {
		fprintf(stderr, "could not mmap %s\n", filename);
		return -1;
	}
	*data = (const uint8_t*)mem;
	*data_size = (size_t)st.st_size;
	return 0;
}

int unmap(const uint8_t* data, size_t data_size)
{
	return munmap((void*)data, data_size);
}
