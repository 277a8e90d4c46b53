/* megalania: drop-in CLI of the reference (src/main.c) over the B200 annealing engine.
 *
 *   megalania [options] <file>  ->  .lzma on stdout, progress on stderr, exit 255 on usage/IO error
 *
 * The structure of src/main.c:64-119 is kept: `steps` rounds of `epochs` restarts, step 0 from the
 * all-literal slab and later steps from the best slab so far, the reference's acceptance rule,
 * then one range-coder pass over the winning slab through the OutputInterface plug-in.  What
 * changes is the engine: every epoch runs on thousands of chains at once (mg_anneal_run), and
 * the iteration budget is explicit because the stock 3 x 200 x n schedule is O(n^2).
 *
 * --rounds N switches to the cooperative search (mg_anneal_merge_regions): all chains work on ONE
 * slab, each confined to a byte region for --round-ms of SM time, the regions' winners are stitched,
 * repaired, priced exactly and kept if cheaper; region boundaries move every round.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host_io.h"
#include "megalania_cuda.h"

static void usage(const char* argv0)
{
	fprintf(stderr,
	        "usage: %s [--chains N] [--iters N] [--epochs N] [--steps N] [--seed N] [--device N] [--top-k N]\n"
	        "          [--rounds N | --time SECONDS] [--round-ms N] [--group N] [--temp T] filename\n",
	        argv0);
}

static int die(const char* what)
{
	fprintf(stderr, "%s: %s\n", what, mg_last_error());
	return -1;
}

int main(int argc, char** argv)
{
	unsigned chains = 1184, iters = 0, epochs = 2, steps = 3, device = 0, top_k = 20;
	unsigned rounds = 0, round_ms = 250, group = 8, temp0 = 4096; /* --temp: 1/2048 bit, cooled linearly to 0 */
	unsigned time_s = 0;                                          /* --time: annealing budget in seconds of SM time */
	unsigned long long seed = 1673551; /* src/main.c:68 */
	const char* filename = NULL;
	for (int i = 1; i < argc; i++) {
		const char* a = argv[i];
		unsigned* target = NULL;
		if (!strcmp(a, "--chains")) target = &chains;
		else if (!strcmp(a, "--iters")) target = &iters;
		else if (!strcmp(a, "--epochs")) target = &epochs;
		else if (!strcmp(a, "--steps")) target = &steps;
		else if (!strcmp(a, "--device")) target = &device;
		else if (!strcmp(a, "--top-k")) target = &top_k;
		else if (!strcmp(a, "--rounds")) target = &rounds;
		else if (!strcmp(a, "--round-ms")) target = &round_ms;
		else if (!strcmp(a, "--group")) target = &group;
		else if (!strcmp(a, "--temp")) target = &temp0;
		else if (!strcmp(a, "--time")) target = &time_s;
		if (target) {
			if (++i >= argc) { usage(argv[0]); return -1; }
			*target = (unsigned)strtoul(argv[i], NULL, 10);
		} else if (!strcmp(a, "--seed")) {
			if (++i >= argc) { usage(argv[0]); return -1; }
			seed = strtoull(argv[i], NULL, 10);
		} else if (a[0] == '-' && a[1] == '-') {
			usage(argv[0]);
			return -1;
		} else if (filename == NULL) {
			filename = a;
		} else {
			usage(argv[0]);
			return -1;
		}
	}
	if (time_s > 0 && round_ms > 0) rounds = (time_s * 1000u + round_ms - 1) / round_ms;
	if (filename == NULL || chains == 0 || (rounds == 0 && (steps == 0 || epochs == 0))) {
		usage(argv[0]);
		return -1;
	}

	InputFile input;
	if (input_file_open(&input, filename) < 0) return -1;
	const uint8_t* file_data = input.data;
	const size_t file_size = input.size;
	if (iters == 0) iters = file_size < 2000 ? (unsigned)file_size : 2000;

	LZMAProperties properties = { 0, 0, 0 };
	mg_ctx* ctx = NULL;
	if (mg_ctx_create(file_data, file_size, properties, (int)device, &ctx)) return die("mg_ctx_create");

	mg_anneal_params params;
	memset(&params, 0, sizeof(params));
	params.chains = chains;
	params.top_k = top_k;
	params.track_best = 1;
	params.seed = seed;
	mg_anneal* an = NULL;
	if (mg_anneal_create(ctx, &params, &an)) return die("mg_anneal_create");

	LZMAPacket* packets_best = malloc(sizeof(LZMAPacket) * file_size);
	uint64_t* best_costs = malloc(sizeof(uint64_t) * chains);
	if (!packets_best || !best_costs) {
		fprintf(stderr, "out of memory\n");
		return -1;
	}
	uint64_t best_perplexity = 0;
	int have_best = 0;
	unsigned long long evals = 0;
	double device_ms = 0;

	if (rounds > 0) {
		/* ---- cooperative regions (megalania_b200/cooperative.py is the same logic) ---------------- */
		if (group == 0 || group > chains) group = chains;
		unsigned nreg = chains / group;
		if (nreg > file_size / 64) nreg = (unsigned)(file_size / 64);
		if (nreg == 0) nreg = 1;
		uint32_t* bounds = malloc(sizeof(uint32_t) * (nreg + 1));
		uint32_t* owners = malloc(sizeof(uint32_t) * nreg);
		uint32_t* regions = malloc(sizeof(uint32_t) * 2 * chains);
		float* temps = calloc(chains, sizeof(float));
		uint64_t* cur = malloc(sizeof(uint64_t) * chains);
		if (!bounds || !owners || !regions || !temps || !cur) {
			fprintf(stderr, "out of memory\n");
			return -1;
		}
		if (mg_anneal_set_slab(an, 0, chains, NULL, 1, 1)) return die("mg_anneal_set_slab");
		unsigned long long lcg = seed;
		for (unsigned r = 0; r < rounds; r++) {
			/* equal regions, boundaries rotated by a pseudo-random shift */
			lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
			const double size = (double)file_size / nreg;
			const double off = (double)((lcg >> 33) % (unsigned long long)(size < 1 ? 1 : size));
			unsigned nb = 0;
			bounds[nb++] = 0;
			for (unsigned k = 1; k < nreg; k++) {
				uint32_t cut = (uint32_t)(off + k * size);
				if (cut >= file_size) cut = (uint32_t)file_size - 1;
				if (cut > bounds[nb - 1]) bounds[nb++] = cut;
			}
			bounds[nb] = (uint32_t)file_size;
			const unsigned regs = nb;
			for (unsigned c = 0; c < chains; c++) {
				regions[2 * c] = bounds[c % regs];
				regions[2 * c + 1] = bounds[c % regs + 1];
			}
			mg_anneal_run_params run;
			memset(&run, 0, sizeof(run));
			run.evals = 1000000;
			run.schedule = MG_SCHEDULE_TEMPERATURE;
			for (unsigned c = 0; c < chains; c++) temps[c] = (float)temp0 * (1.f - (float)(r + 1) / (float)rounds);
			run.temperatures = temps; /* uphill moves with probability exp(-delta / T); the last round is T = 0 */
			run.first_eval = MG_CONTINUE_EVALS;
			run.cycle_budget = (uint64_t)round_ms * 1965000ull;
			run.regions = regions;
			mg_anneal_stats stats;
			if (mg_anneal_run(an, &run, &stats)) return die("mg_anneal_run");
			evals += stats.evals;
			device_ms += stats.kernel_ms;
			if (mg_anneal_costs(an, cur, NULL)) return die("mg_anneal_costs");
			unsigned best_chain = 0, worst_chain = 0;
			for (unsigned k = 0; k < regs; k++) owners[k] = k;
			for (unsigned c = 0; c < chains; c++) {
				if (cur[c] < cur[owners[c % regs]]) owners[c % regs] = c;
				if (cur[c] < cur[best_chain]) best_chain = c;
				if (cur[c] > cur[worst_chain]) worst_chain = c;
			}
			if (worst_chain == best_chain) worst_chain = (best_chain + 1) % chains;
			uint64_t merged = 0;
			unsigned winner = best_chain;
			if (chains > 1) {
				if (mg_anneal_merge_regions(an, regs, bounds, owners, worst_chain, &merged)) return die("mg_anneal_merge_regions");
				if (merged <= cur[best_chain]) winner = worst_chain;
			}
			best_perplexity = winner == best_chain ? cur[best_chain] : merged;
			if (mg_anneal_broadcast_chain(an, winner)) return die("mg_anneal_broadcast_chain");
			fprintf(stderr, "current file size: %f\tround: %u/%u - %s, %llu evals, %.0f evals/s\n", 18 + best_perplexity / 16384.f,
			        r + 1, rounds, winner == best_chain ? "best chain" : "merged", evals, evals / (device_ms / 1e3));
		}
		if (mg_anneal_get_slab(an, 0, 0, packets_best)) return die("mg_anneal_get_slab");
		have_best = 1;
		free(bounds);
		free(owners);
		free(regions);
		free(temps);
		free(cur);
		steps = 0;
	}

	for (unsigned step = 0; step < steps; step++) {
		for (unsigned epoch = 0; epoch < epochs; epoch++) {
			/* src/main.c:71-77: fresh slab in step 0, the best so far afterwards; cost 0 means
			 * "first proposal always accepted" */
			if (mg_anneal_set_slab(an, 0, chains, (step != 0 && have_best) ? packets_best : NULL, 0, 0))
				return die("mg_anneal_set_slab");
			mg_anneal_run_params run;
			memset(&run, 0, sizeof(run));
			run.evals = iters;
			run.schedule = MG_SCHEDULE_REFERENCE;
			run.step = step;
			run.num_iters = iters;
			mg_anneal_stats stats;
			if (mg_anneal_run(an, &run, &stats)) return die("mg_anneal_run");
			evals += stats.evals;
			device_ms += stats.kernel_ms;
			if (mg_anneal_costs(an, NULL, best_costs)) return die("mg_anneal_costs");
			unsigned arg = 0;
			for (unsigned c = 1; c < chains; c++)
				if (best_costs[c] != 0 && (best_costs[arg] == 0 || best_costs[c] < best_costs[arg])) arg = c;
			if (best_costs[arg] != 0 && (!have_best || best_costs[arg] < best_perplexity)) {
				best_perplexity = best_costs[arg];
				if (mg_anneal_get_slab(an, arg, 1, packets_best)) return die("mg_anneal_get_slab");
				have_best = 1;
			}
			fprintf(stderr, "current file size: %f\tstep: %u\tepoch: %04u - %llu evals, %.0f evals/s\n",
			        18 + best_perplexity / 16384.f, step + 1, epoch, evals, evals / (device_ms / 1e3));
		}
	}
	mg_anneal_destroy(an);
	if (!have_best) {
		/* no position had an alternative packet (the reference loops forever here): all literals */
		for (size_t i = 0; i < file_size; i++) {
			packets_best[i].type = LITERAL;
			packets_best[i].dist = 0;
			packets_best[i].len = 1;
		}
	}

	OutputInterface output;
	StreamSink sink;
	stream_sink_init(&output, &sink, stdout);
	if (mg_encode_slab(ctx, packets_best, &output)) return die("mg_encode_slab");
	if (fflush(stdout) != 0 || sink.failed) {
		fprintf(stderr, "could not write the output stream\n");
		return -1;
	}

	free(best_costs);
	free(packets_best);
	mg_ctx_destroy(ctx);
	input_file_close(&input);
	return 0;
}
