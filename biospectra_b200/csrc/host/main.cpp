// biospectra_gpu -- command-line driver mirroring biospectra.BioSpectra.main for the two
// modes on the hot path (src/biospectra/BioSpectra.java:56-108):
//   i | index       -j conf.json | -i <indexdir>   -r <reference dir>      (CommandArgumentIndex.java:29-36)
//   lc | lclassify  -j conf.json | -i <indexdir>   -in <reads> -out <dir>  (CommandArgumentLocalClassifier.java:29-39)
// rc/svr (RabbitMQ) and simulate stay with the reference.
#include "../../../include/biospectra_gpu.h"

#include <string>
#include <cstdlib>

#include <cstdio>
#include <cstring>
#include <strings.h>

static void print_help() {
    printf("usage: biospectra_gpu i|index [-j conf.json | -i indexdir] -r refdir [--device N]\n"
           "       biospectra_gpu lc|lclassify [-j conf.json | -i indexdir] -in reads -out outdir [--device N]\n"
           "                 [--per-read THREADS]   one Classifier.classify call per read from THREADS pool threads, like the\n"
           "                                        reference's LocalClassifier (0 = worker_threads of the configuration)\n");
}

int main(int argc, char** argv) {
    if (argc < 2) { print_help(); return 0; }
    std::string mode = argv[1], json, index, ref, in, out;
    int device = 0, per_read = -1;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](std::string& dst) { if (i + 1 < argc) dst = argv[++i]; };
        if (a == "-j" || a == "--json") val(json);
        else if (a == "-i" || a == "--index") val(index);
        else if (a == "-r" || a == "--ref") val(ref);
        else if (a == "-in" || a == "--input") val(in);
        else if (a == "-out" || a == "--output") val(out);
        else if (a == "--device") { std::string d; val(d); device = atoi(d.c_str()); }
        else if (a == "--per-read") { std::string d; val(d); per_read = atoi(d.c_str()); }
        else if (a == "-h" || a == "--help") { print_help(); return 0; }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); print_help(); return 1; }
    }
    int rc = 0;
    const char* j = json.empty() ? nullptr : json.c_str();
    const char* ix = index.empty() ? nullptr : index.c_str();
    if (!strcasecmp(mode.c_str(), "i") || !strcasecmp(mode.c_str(), "index")) {
        printf("Indexing...\n");
        rc = bsp_host_index(j, ix, ref.c_str(), device);
    } else if (!strcasecmp(mode.c_str(), "lc") || !strcasecmp(mode.c_str(), "lclassify")) {
        printf("Classifying (LOCAL mode)...\n");
        rc = per_read >= 0 ? bsp_host_classify_local_per_read(j, ix, in.c_str(), out.c_str(), device, per_read)
                           : bsp_host_classify_local(j, ix, in.c_str(), out.c_str(), device);
    } else { print_help(); }
    if (rc != 0) { fprintf(stderr, "Exception: %s\n", bsp_host_last_error()); return 2; }
    return 0;
}
