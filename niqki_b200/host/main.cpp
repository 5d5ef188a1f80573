// main.cpp — `niqki_b200`: the niqki command line (/root/reference/src/niqki.cpp:102-185 option
// table, :229-456 main) on top of the B200 engine.  Same flags, defaults, phase order, working-
// directory behaviour (-I/-i/-M chdir to the list's directory, -Q/-l do not) and the
// same "Informations" table on stdout.  Behaviour follows the code, not the README, where they
// differ: --minjac defaults to 0, --querylines' short flag is -l, -P is a no-op
// and the output is text.  Additive flags: --device N, --gpus N, --binary, --nowrap, --threads N, --verbose.
#include <fcntl.h>
#include <libgen.h>
#include <limits.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "engine.hpp"

namespace {

enum ArgKind { kNone, kNonEmpty, kNumeric };
struct Flag {
  const char* id;
  const char* shortname;
  const char* longname;
  ArgKind arg;
  const char* help;
};

const Flag kFlags[] = {
    {"index", "I", "index", kNonEmpty, "  --index, -I <filename> \tInput file of files to Index."},
    {"query", "Q", "query", kNonEmpty, "  --query, -Q <filename> \tInput file of file to Query."},
    {"indexlines", "i", "indexlines", kNonEmpty, "  --indexlines, -i <filename> \tfa/fq file where each record is a separate entry to Index."},
    {"querylines", "l", "querylines", kNonEmpty, "  --querylines, -l <filename> \tfa/fq file where each record is a separate entry to Query."},
    {"kmer", "K", "kmer", kNumeric, "  --kmer, -K <int> \tKmer size (31)."},
    {"sketch", "S", "sketch", kNumeric, "  --sketch, -S <int> \tSet sketch size to 2^S (15)."},
    {"output", "O", "output", kNonEmpty, "  --output, -O <filename> \tOutput file (niqkiOutput.gz)"},
    {"minjac", "J", "minjac", kNonEmpty, "  --minjac, -J <float> \tMinimal jaccard Index to report (0)."},
    {"pretty", "P", "pretty", kNone, "  --pretty, -P \tHuman-readable outfile (always on, as in the reference binary)."},
    {"matrix", "M", "matrix", kNonEmpty, "  --matrix, -M <filename> \tOutput the all-vs-all matrix of the given file of files."},
    {"word", "W", "word", kNumeric, "  --word, -W <int> \tFingerprint size (12)."},
    {"gsize", "G", "Genomes_sizes", kNumeric, "  --Genomes_sizes, -G <int> \tRough expectation of the genome sizes."},
    {"hhl", "H", "HHL", kNumeric, "  --HHL, -H <int> \tSize of the hyperloglog section (4)."},
    {"dump", "D", "dump", kNonEmpty, "  --dump, -D <filename> \tDump the current index to the given file."},
    {"load", "L", "load", kNonEmpty, "  --load, -L <filename> \tLoad an index from the given file."},
    {"download", "Iddl", "indexdownload", kNonEmpty, "  --indexdownload, -Iddl <filename> \t(not available: this build has no network path)"},
    {"logo", "", "logo", kNone, "  --logo \tPrint ASCII art logo, then exit."},
    {"help", "h", "help", kNone, "  --help, -h \tPrint usage and exit."},
    // additive
    {"device", "", "device", kNumeric, "  --device <int> \tFirst CUDA device to run on (0)."},
    {"gpus", "", "gpus", kNumeric, "  --gpus <int> \tShard the index by genome id over this many devices (1); queries are all-gathered over NCCL."},
    {"binary", "", "binary", kNone, "  --binary \tWrite query results in the reference's binary record format."},
    {"nowrap", "", "nowrap", kNone, "  --nowrap \tMatrix counters keep 32 bits (the reference wraps at 65536 for S >= 16)."},
    {"threads", "", "threads", kNumeric, "  --threads <int> \tReader threads for file-of-files ingest."},
    {"verbose", "", "verbose", kNone, "  --verbose \tNotes on stderr."},
};

void print_usage(std::ostream& os) {
  os << "niqki_b200 — NIQKI on B200 (" << nq_version() << ")\n";
  for (const Flag& f : kFlags) os << f.help << "\n";
}

bool is_numeric(const char* s) {
  char* end = nullptr;
  strtol(s, &end, 10);
  return end != s && *end == 0;
}

int old_wd = -1;
void change_dir_from_filename(const char* fname) {  // niqki.cpp:199-214
  old_wd = open(".", O_CLOEXEC);
  char copy[PATH_MAX];
  strncpy(copy, fname, PATH_MAX);
  copy[PATH_MAX - 1] = '\0';
  errno = 0;
  if (chdir(dirname(copy))) std::cout << "Error: " << strerror(errno) << std::endl;
}
void restore_dir() {  // niqki.cpp:218-224
  errno = 0;
  if (fchdir(old_wd)) std::cout << "Error: " << strerror(errno) << std::endl;
  close(old_wd);
}
std::string base_name(const std::string& p) { return p.substr(p.find_last_of("/\\") + 1); }

void row(const char* label, double v) {
  std::cout << label << std::setw(30) << std::setfill(' ') << v << " |" << std::endl;
}

}  // namespace

int main(int argc, char** argv) {
  using clock = std::chrono::system_clock;
  std::map<std::string, std::string> opt;
  std::vector<std::string> stray;
  bool bad = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    const Flag* hit = nullptr;
    std::string value;
    bool has_value = false;
    if (a.rfind("--", 0) == 0) {
      std::string name = a.substr(2);
      const size_t eq = name.find('=');
      if (eq != std::string::npos) {
        value = name.substr(eq + 1);
        name = name.substr(0, eq);
        has_value = true;
      }
      for (const Flag& f : kFlags)
        if (name == f.longname) hit = &f;
    } else if (a.size() >= 2 && a[0] == '-') {
      for (const Flag& f : kFlags) {
        const size_t n = strlen(f.shortname);
        if (n && a.compare(1, std::string::npos, f.shortname) == 0) hit = &f;
      }
      if (!hit)  // -Svalue form
        for (const Flag& f : kFlags)
          if (strlen(f.shortname) == 1 && a[1] == f.shortname[0] && f.arg != kNone) {
            hit = &f;
            value = a.substr(2);
            has_value = true;
          }
    } else {
      stray.push_back(a);
      continue;
    }
    if (!hit) {
      std::cerr << "Unknown option '" << a << "'\n";
      bad = true;
      continue;
    }
    if (hit->arg != kNone) {
      if (!has_value) {
        if (i + 1 >= argc) {
          std::cerr << "Option '" << a << "' requires a non-empty argument\n";
          bad = true;
          continue;
        }
        value = argv[++i];
      }
      if (value.empty() || (hit->arg == kNumeric && !is_numeric(value.c_str()))) {
        std::cerr << "Option '" << a << "' requires a " << (hit->arg == kNumeric ? "numeric" : "non-empty") << " argument\n";
        bad = true;
        continue;
      }
    }
    opt[hit->id] = value;  // the last occurrence wins, as options[X].last() does
  }
  if (bad) {
    std::cout << "Bad usage!!!" << std::endl;
    return EXIT_FAILURE;
  }
  if (opt.count("help") || argc <= 1) {
    print_usage(std::clog);
    return EXIT_SUCCESS;
  }
  auto has = [&](const char* k) { return opt.count(k) != 0; };
  const int K = has("kmer") ? atoi(opt["kmer"].c_str()) : 31;
  const int S = has("sketch") ? atoi(opt["sketch"].c_str()) : 15;
  const int H = has("hhl") ? atoi(opt["hhl"].c_str()) : 4;
  const int W = has("word") ? atoi(opt["word"].c_str()) : 12;
  const double min_fract = has("minjac") ? atof(opt["minjac"].c_str()) : 0;
  const unsigned genomes_sizes = has("gsize") ? (unsigned)atoi(opt["gsize"].c_str()) : 0;
  for (size_t i = 0; i < stray.size(); ++i) {
    std::cout << "Non-option argument #" << i << " is " << stray[i] << std::endl;
    std::cout << "Ignoring unknown argument '" << stray[i] << "'" << std::endl;
  }
  if (!stray.empty()) {
    std::cout << "Bad usage!!!" << std::endl;
    return EXIT_FAILURE;
  }
  const std::string out_file = has("output") ? opt["output"] : "niqkiOutput.gz";
  if (has("download")) {
    std::cout << "--indexdownload is not available in niqki_b200 (no network path)" << std::endl;
    return EXIT_FAILURE;
  }

  std::cout << "+-------------------------------------------------------------------+" << std::endl;
  std::cout << "|                            Informations                           |" << std::endl;
  std::cout << "+-----------------------------------+-------------------------------+" << std::endl;
  nqh::EngineOptions eo;
  eo.device = has("device") ? atoi(opt["device"].c_str()) : 0;
  eo.gpus = has("gpus") ? std::max(1, atoi(opt["gpus"].c_str())) : 1;
  eo.binary_output = has("binary");
  eo.matrix_nowrap = has("nowrap");
  eo.reader_threads = has("threads") ? (unsigned)atoi(opt["threads"].c_str()) : 0;
  eo.verbose = has("verbose");
  std::unique_ptr<nqh::Engine> index;
  try {
    if (has("load")) index.reset(new nqh::Engine(opt["load"], out_file, eo));
    else index.reset(new nqh::Engine((uint32_t)S, (uint32_t)K, (uint32_t)W, (uint32_t)H, out_file, min_fract, eo));
    if (genomes_sizes != 0) index->select_best_H(genomes_sizes);

    auto start = clock::now();
    auto warn_unreadable = [](const std::string& f) {
      std::ifstream ifs(f);
      if (!ifs) std::cout << "Unable to open the file '" << f << "'" << std::endl;
    };
    if (has("index")) {
      const std::string list_file = opt["index"];
      warn_unreadable(list_file);
      change_dir_from_filename(list_file.c_str());
      index->insert_file_of_file_whole(base_name(list_file));
      restore_dir();
    }
    if (has("indexlines")) {
      const std::string list_file = opt["indexlines"];
      warn_unreadable(list_file);
      change_dir_from_filename(list_file.c_str());
      index->insert_file_lines(base_name(list_file));
      restore_dir();
    }
    if (has("dump")) index->dump_index_disk(opt["dump"]);
    auto endindex = clock::now();
    std::chrono::duration<double> elapsed = endindex - start;
    row("| Indexing lasted (s)               |", elapsed.count());

    if (has("matrix")) {
      const std::string matrix_file = opt["matrix"];
      warn_unreadable(matrix_file);
      if (!has("index") && !has("indexlines")) {
        start = clock::now();
        change_dir_from_filename(matrix_file.c_str());
        index->insert_file_of_file_whole(base_name(matrix_file));
        restore_dir();
        endindex = clock::now();
        elapsed = endindex - start;
        row("| Indexing lasted (s)               |", elapsed.count());
      }
      change_dir_from_filename(matrix_file.c_str());
      start = clock::now();
      index->query_matrix();
      elapsed = clock::now() - start;
      row("| Query lasted (s)                  |", elapsed.count());
      restore_dir();
    }
    if (has("query")) {
      warn_unreadable(opt["query"]);
      index->query_file_of_file_whole(opt["query"]);
    }
    if (has("querylines")) {
      warn_unreadable(opt["querylines"]);
      index->query_file_lines(opt["querylines"]);
    }
    index->close_output();
    auto end = clock::now();
    elapsed = end - endindex;
    row("| Query lasted (s)                  |", elapsed.count());
    elapsed = end - start;
    row("| Whole run lasted (s)              |", elapsed.count());

    if (has("logo")) {
      std::ifstream logo("../resources/niqki.ascii");
      std::string line;
      if (logo.is_open())
        while (std::getline(logo, line)) std::cout << line << '\n';
      else
        std::cout << "Unable to open file :'../resources/niqki.ascii'" << std::endl;
      return EXIT_SUCCESS;
    }
    std::cout << "+-----------------------------------+-------------------------------+" << std::endl;
    std::cout << "| k-mer size                        |" << std::setw(30) << std::setfill(' ') << K << " |" << std::endl
              << "| S                                 |" << std::setw(30) << std::setfill(' ') << S << " |" << std::endl
              << "| Number of fingerprints            |" << std::setw(30) << std::setfill(' ') << index->params().F << " |" << std::endl
              << "| W                                 |" << std::setw(30) << std::setfill(' ') << W << " |" << std::endl
              << "| H                                 |" << std::setw(30) << std::setfill(' ') << H << " |" << std::endl
              << "| Number of indexed genomes         |" << std::setw(30) << std::setfill(' ') << index->getNbGenomes() << " |" << std::endl;
    std::cout << "| GPU kernel launches               |" << std::setw(30) << std::setfill(' ') << index->kernel_launches() << " |" << std::endl;
    std::cout << "+-----------------------------------+-------------------------------+" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "niqki_b200: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
