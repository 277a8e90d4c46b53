#include "file_output.h"

static bool stream_write(OutputInterface* output, const void* data, size_t data_size)
{
	FILE* stream = (FILE*)output->private_data;
	return fwrite(data, 1, data_size, stream) == data_size;
}

void file_output_new(OutputInterface* output, FILE* file)
{
	output->write = stream_write;
	output->private_data = file;
}
