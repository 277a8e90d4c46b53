#include "host_io.h"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* OutputInterface.write: all of `data` or failure.  stdio may accept a buffer in pieces (a pipe
 * that fills up, a signal): carry on from where it stopped instead of reporting a short write. */
static bool stream_sink_write(OutputInterface* output, const void* data, size_t data_size)
{
	StreamSink* sink = (StreamSink*)output->private_data;
	const unsigned char* cursor = (const unsigned char*)data;
	if (sink->failed) return false;
	while (data_size > 0) {
		size_t done = fwrite(cursor, 1, data_size, sink->stream);
		if (done == 0) {
			if (ferror(sink->stream) && errno == EINTR) {
				clearerr(sink->stream);
				continue;
			}
			sink->failed = 1;
			return false;
		}
		cursor += done;
		data_size -= done;
		sink->bytes_written += done;
	}
	return true;
}

void stream_sink_init(OutputInterface* output, StreamSink* sink, FILE* stream)
{
	sink->stream = stream;
	sink->bytes_written = 0;
	sink->failed = 0;
	output->write = stream_sink_write;
	output->private_data = sink;
}

int input_file_open(InputFile* in, const char* path)
{
	struct stat info;
	in->data = NULL;
	in->size = 0;
	int fd = open(path, O_RDONLY);
	if (fd < 0 || fstat(fd, &info) != 0) {
		fprintf(stderr, "%s: %s\n", path, strerror(errno));
		if (fd >= 0) close(fd);
		return -1;
	}
	void* view = info.st_size > 0 ? mmap(NULL, (size_t)info.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : MAP_FAILED;
	close(fd); /* the mapping keeps the file alive */
	if (view == MAP_FAILED) {
		fprintf(stderr, "could not mmap %s\n", path);
		return -1;
	}
	in->data = (const uint8_t*)view;
	in->size = (size_t)info.st_size;
	return 0;
}

void input_file_close(InputFile* in)
{
	if (in->data) munmap((void*)in->data, in->size);
	in->data = NULL;
	in->size = 0;
}
