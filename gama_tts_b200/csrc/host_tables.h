#ifndef GTTS_HOST_TABLES_H_
#define GTTS_HOST_TABLES_H_

#include <cstdint>
#include <vector>

#include "../../include/gtts_b200.h"
#include "tube_types.h"

namespace gtts {

std::vector<double> designGlottalFir(double beta = 0.2, double gamma = 0.1, double cutoff = 0.00000001);
void buildSrcTables(double* h, double* dh);
int internalRate(const gtts_voice_config& c);
int controlSteps(int fs, double controlRate);
// Returns nullptr on success, else a static error text.
const char* deriveVoice(const gtts_voice_config& c, VoiceDev& v);
int64_t outputLength(const VoiceDev& v, int64_t nInternal);
void buildWavetable(const VoiceDev& v, double* table512);
void shardPlan(const int64_t* cost, int64_t n, int shards, int32_t* shardOf);

} // namespace gtts
#endif
