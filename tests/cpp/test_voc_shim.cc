// Exercises the ORBVocabulary shim: loadFromTextFile (DBoW2 text format) + transform; prints per-feature node ids so that
// the Python side can compare them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBVocabulary.h"

int main(int argc, char** argv) {
  if (argc < 5) return 2;  // voc.txt features.bin n out.txt
  ORB_SLAM2::ORBVocabulary voc;
  if (!voc.loadFromTextFile(argv[1])) return 3;
  const int n = atoi(argv[3]);
  std::vector<uint8_t> feats((size_t)n * 32);
  FILE* f = fopen(argv[2], "rb");
  if (!f || fread(feats.data(), 32, n, f) != (size_t)n) return 4;
  fclose(f);
  ORB_SLAM2::BowVector bow;
  ORB_SLAM2::FeatureVector fv;
  voc.transform(feats.data(), n, bow, fv, 2);
  FILE* o = fopen(argv[4], "w");
  fprintf(o, "%u %zu %zu\n", voc.size(), bow.size(), fv.size());
  for (auto& kv : bow) fprintf(o, "w %u %.17g\n", kv.first, kv.second);
  for (auto& kv : fv) {
    fprintf(o, "n %u", kv.first);
    for (unsigned i : kv.second) fprintf(o, " %u", i);
    fprintf(o, "\n");
  }
  fclose(o);
  return 0;
}
