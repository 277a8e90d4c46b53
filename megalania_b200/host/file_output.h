/* OutputInterface over a stdio stream (the role of the reference's src/file_output.{c,h}). */
#ifndef MEGALANIA_FILE_OUTPUT_H
#define MEGALANIA_FILE_OUTPUT_H
#include <stdio.h>

#include "output_interface.h"

void file_output_new(OutputInterface* output, FILE* file);

#endif
