// currennt_b200 -- command line trainer / forward-pass tool with the reference's option names and console output
// (currennt/src/main.cpp:98-490, currennt_lib/src/Configuration.cpp:110-330).  The compute path is the B200 library only:
// there is no CPU mode (--cuda false is an error).
//
// Data parallelism (not in the reference): start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
// MASTER_PORT in the environment (torchrun's variables).  --parallel_sequences is per process.  Rank 0 prints and saves.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "NetCdf.hpp"
#include "NeuralNetwork.hpp"
#include "Optimizer.hpp"

namespace {

// ---------------------------------------------------------------- options (Configuration.cpp:110-330)
struct Options {
    std::map<std::string, std::string> v;
    static const std::map<std::string, std::string> &defaults()
    {
        static const std::map<std::string, std::string> d = {
            {"network", "network.jsn"}, {"cuda", "true"}, {"list_devices", "false"}, {"parallel_sequences", "1"}, {"random_seed", "0"},
            {"ff_output_format", "single_csv"}, {"ff_output_file", "ff_output.csv"}, {"ff_output_kind", "9"}, {"feature_period", "10"},
            {"ff_input_file", ""}, {"revert_std", "true"},
            {"train", "false"}, {"stochastic", "false"}, {"hybrid_online_batch", "false"}, {"shuffle_fractions", "false"},
            {"shuffle_sequences", "false"}, {"max_epochs", "-1"}, {"max_epochs_no_best", "20"}, {"validate_every", "1"}, {"test_every", "1"},
            {"optimizer", "steepest_descent"}, {"learning_rate", "1e-5"}, {"momentum", "0.9"}, {"weight_noise_sigma", "0"},
            {"save_network", "trained_network.jsn"},
            {"autosave", "false"}, {"autosave_best", "false"}, {"autosave_prefix", ""}, {"continue", ""},
            {"train_file", ""}, {"val_file", ""}, {"test_file", ""}, {"train_fraction", "1"}, {"val_fraction", "1"}, {"test_fraction", "1"},
            {"truncate_seq", "0"}, {"input_noise_sigma", "0"}, {"input_left_context", "0"}, {"input_right_context", "0"},
            {"output_time_lag", "0"}, {"cache_path", ""},
            {"weights_dist", "uniform"}, {"weights_uniform_min", "-0.1"}, {"weights_uniform_max", "0.1"},
            {"weights_normal_sigma", "0.1"}, {"weights_normal_mean", "0"},
            {"gemm_mode", "strict"},          // B200 only: strict = fp32-accurate (3xTF32), fast = 1xTF32 forward projections
        };
        return d;
    }
    static std::string trim(const std::string &s)
    {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    void set(const std::string &key, const std::string &value, bool override)
    {
        if (key != "options_file" && key != "help" && !defaults().count(key)) throw std::runtime_error("unrecognised option '" + key + "'");
        if (override || !v.count(key)) v[key] = value;
    }
    void parseFile(const std::string &path)
    {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("Could not open options file '" + path + "'");
        std::string line;
        while (std::getline(f, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            line = trim(line);
            if (line.empty()) continue;
            const size_t eq = line.find('=');
            if (eq == std::string::npos) throw std::runtime_error("Error while parsing the options file: '" + line + "'");
            set(trim(line.substr(0, eq)), trim(line.substr(eq + 1)), false);      // the command line wins (boost::program_options: first store wins)
        }
    }
    void parse(int argc, const char **argv)
    {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (a.rfind("--", 0) != 0) { set("options_file", a, true); continue; }       // positional = options file (Configuration.cpp:197-198)
            a = a.substr(2);
            const size_t eq = a.find('=');
            if (eq != std::string::npos) set(a.substr(0, eq), a.substr(eq + 1), true);
            else if (a == "help") set("help", "true", true);
            else { if (i + 1 >= argc) throw std::runtime_error("missing value for --" + a); set(a, argv[++i], true); }
        }
        if (v.count("options_file")) parseFile(v["options_file"]);
        for (const auto &d : defaults()) if (!v.count(d.first)) v[d.first] = d.second;
    }
    // "key=value;;;..." as stored in autosave files (Configuration.cpp:46-66)
    std::string serialize() const
    {
        std::string s;
        for (const auto &kv : v) if (kv.first != "continue" && kv.first != "options_file" && kv.first != "help") s += kv.first + '=' + kv.second + ";;;";
        return s;
    }
    // --continue: every option is taken from the autosave file's "configuration" string (Configuration.cpp:68-100, 239-250)
    void restoreFrom(const std::string &serialized)
    {
        const std::string cont = v.count("continue") ? v["continue"] : "";
        v.clear();
        size_t pos = 0;
        while (pos < serialized.size()) {
            size_t end = serialized.find(";;;", pos);
            if (end == std::string::npos) end = serialized.size();
            const std::string item = serialized.substr(pos, end - pos);
            const size_t eq = item.find('=');
            if (eq != std::string::npos) set(trim(item.substr(0, eq)), trim(item.substr(eq + 1)), true);
            pos = end + 3;
        }
        for (const auto &d : defaults()) if (!v.count(d.first)) v[d.first] = d.second;
        v["continue"] = cont;
    }
    const std::string &str(const std::string &k) const { return v.at(k); }
    double num(const std::string &k) const
    {
        char *end = nullptr; const std::string &s = v.at(k); const double d = std::strtod(s.c_str(), &end);
        if (end == s.c_str() || *end) throw std::runtime_error("invalid value '" + s + "' for --" + k);
        return d;
    }
    bool flag(const std::string &k) const
    {
        std::string s = v.at(k); for (char &c : s) c = (char)std::tolower(c);
        if (s == "true" || s == "1" || s == "on" || s == "yes") return true;
        if (s == "false" || s == "0" || s == "off" || s == "no") return false;
        throw std::runtime_error("invalid value '" + v.at(k) + "' for --" + k);
    }
    std::vector<std::string> list(const std::string &k) const
    {
        std::vector<std::string> out; std::stringstream ss(v.at(k)); std::string item;
        while (std::getline(ss, item, ',')) { item = trim(item); if (!item.empty()) out.push_back(item); }
        return out;
    }
};

void printHelp()
{
    std::printf("Usage: currennt_b200 [options] [options-file]\n\nOptions (same names as the reference; defaults in brackets):\n");
    for (const auto &d : Options::defaults()) std::printf("  --%-22s [%s]\n", d.first.c_str(), d.second.c_str());
    std::printf("  --%-22s\n", "options_file");
}

// ---------------------------------------------------------------- rendezvous: rank 0 sends 128 B NCCL id + 4 B seed to every rank
void sendAll(int fd, const void *buf, size_t n)
{
    const char *p = (const char *)buf;
    while (n) { ssize_t k = ::send(fd, p, n, 0); if (k <= 0) throw std::runtime_error("rendezvous: send failed"); p += k; n -= (size_t)k; }
}
void recvAll(int fd, void *buf, size_t n)
{
    char *p = (char *)buf;
    while (n) { ssize_t k = ::recv(fd, p, n, 0); if (k <= 0) throw std::runtime_error("rendezvous: recv failed"); p += k; n -= (size_t)k; }
}
void exchange(int rank, int world, char *blob, size_t bytes)
{
    const char *addr = std::getenv("MASTER_ADDR"); if (!addr) addr = "127.0.0.1";
    const int port = (std::getenv("MASTER_PORT") ? std::atoi(std::getenv("MASTER_PORT")) : 29500) + 17;
    if (rank == 0) {
        int ls = ::socket(AF_INET, SOCK_STREAM, 0); int one = 1;
        ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in sa{}; sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((uint16_t)port);
        if (::bind(ls, (sockaddr *)&sa, sizeof sa) || ::listen(ls, world)) throw std::runtime_error("rendezvous: cannot listen on port " + std::to_string(port));
        for (int i = 1; i < world; ++i) {
            int fd = ::accept(ls, nullptr, nullptr);
            if (fd < 0) throw std::runtime_error("rendezvous: accept failed");
            sendAll(fd, blob, bytes); ::close(fd);
        }
        ::close(ls);
    } else {
        addrinfo hints{}, *res = nullptr; hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
        if (::getaddrinfo(addr, std::to_string(port).c_str(), &hints, &res) || !res) throw std::runtime_error("rendezvous: cannot resolve MASTER_ADDR");
        int fd = -1;
        for (int attempt = 0; attempt < 600; ++attempt) {                          // up to 60 s for rank 0 to come up
            fd = ::socket(AF_INET, SOCK_STREAM, 0);
            if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
            ::close(fd); fd = -1; ::usleep(100000);
        }
        ::freeaddrinfo(res);
        if (fd < 0) throw std::runtime_error("rendezvous: cannot reach rank 0");
        recvAll(fd, blob, bytes); ::close(fd);
    }
}

// ---------------------------------------------------------------- helpers
std::string printfRow(bool echo, const char *format, ...)
{
    char buffer[512];
    va_list args; va_start(args, format); std::vsnprintf(buffer, sizeof buffer, format, args); va_end(args);
    if (echo) { std::fputs(buffer, stdout); std::fflush(stdout); }
    return buffer;
}

void saveNetwork(const NeuralNetwork &nn, const std::string &filename)                  // main.cpp:681-700
{
    helpers::JsonDocument doc = helpers::JsonValue::makeObject();
    nn.exportLayers(doc);
    nn.exportWeights(doc);
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("Cannot open file '" + filename + "' for writing");
    f << doc.serialize(true);
}

std::string replaceAll(std::string s, const std::string &from, const std::string &to)
{
    for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
    return s;
}

// autosave file: configuration + epoch table + network + optimizer state (main.cpp:702-741); name <prefix>_epochNNN.autosave
void saveState(const NeuralNetwork &nn, optimizers::SteepestDescentOptimizer &optimizer, const Options &opt, const std::string &infoRows)
{
    helpers::JsonDocument doc = helpers::JsonValue::makeObject();
    doc.member("configuration") = helpers::JsonValue::makeString(opt.serialize());
    doc.member("info_rows") = helpers::JsonValue::makeString(replaceAll(infoRows, "\n", ";;;"));
    nn.exportLayers(doc);
    nn.exportWeights(doc);
    optimizer.exportState(doc);
    char name[64]; std::snprintf(name, sizeof name, "epoch%03d.autosave", optimizer.currentEpoch());
    const std::string prefix = opt.str("autosave_prefix");
    const std::string path = prefix.empty() ? std::string(name) : prefix + "_" + name;
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("Cannot open file");
    f << doc.serialize(true);
}

void makeDirs(const std::string &path)
{
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') ::mkdir(path.substr(0, i).c_str(), 0777);
}

std::string dirName(const std::string &p) { size_t s = p.find_last_of('/'); return s == std::string::npos ? "" : p.substr(0, s); }
std::string baseName(const std::string &p) { size_t s = p.find_last_of('/'); return s == std::string::npos ? p : p.substr(s + 1); }
std::string relativeDir(const std::string &tag) { std::string d = dirName(tag); while (!d.empty() && d[0] == '/') d.erase(0, 1); return d; }

void put32(std::ostream &f, uint32_t x) { unsigned char b[4] = {(unsigned char)(x >> 24), (unsigned char)(x >> 16), (unsigned char)(x >> 8), (unsigned char)x}; f.write((char *)b, 4); }
void put16(std::ostream &f, uint16_t x) { unsigned char b[2] = {(unsigned char)(x >> 8), (unsigned char)x}; f.write((char *)b, 2); }

struct Ctx {
    bl_ctx *p = nullptr;
    ~Ctx() { if (p) bl_ctx_destroy(p); }
};

int run(const Options &opt)
{
    const int rank = std::getenv("RANK") ? std::atoi(std::getenv("RANK")) : 0;
    const int world = std::getenv("WORLD_SIZE") ? std::atoi(std::getenv("WORLD_SIZE")) : 1;
    const int localRank = std::getenv("LOCAL_RANK") ? std::atoi(std::getenv("LOCAL_RANK")) : rank;
    const bool chief = rank == 0;
    const bool training = opt.flag("train");

    if (!opt.flag("cuda")) throw std::runtime_error("--cuda false: this build has no CPU path");
    if (opt.str("optimizer") != "steepest_descent") throw std::runtime_error("Unknown optimizer type");
    Ctx ctx;
    if (bl_ctx_create(localRank, nullptr, &ctx.p)) throw std::runtime_error(std::string("bl_ctx_create: ") + bl_last_error(nullptr));
    const std::string gm = opt.str("gemm_mode");
    if (gm != "strict" && gm != "fast") throw std::runtime_error("--gemm_mode must be strict or fast");
    bl_ctx_set_gemm_mode(ctx.p, gm == "fast" ? BL_GEMM_FAST : BL_GEMM_STRICT);

    // seed + communicator
    struct { char id[128]; unsigned seed; } blob;
    std::memset(&blob, 0, sizeof blob);
    blob.seed = (unsigned)opt.num("random_seed");
    if (blob.seed == 0) blob.seed = (unsigned)std::time(nullptr) ^ (unsigned)::getpid();      // 0 = auto (Configuration.cpp:236-238)
    bl_comm *comm = nullptr;
    if (world > 1) {
        if (chief && bl_comm_unique_id(blob.id)) throw std::runtime_error(bl_last_error(ctx.p));
        exchange(rank, world, (char *)&blob, sizeof blob);
        if (bl_comm_create(ctx.p, rank, world, blob.id, &comm)) throw std::runtime_error(bl_last_error(ctx.p));
    }
    Configuration &cfg = Configuration::instance();
    cfg.randomSeed = blob.seed;
    if (opt.str("weights_dist") == "uniform") cfg.weightsUniform = true;
    else if (opt.str("weights_dist") == "normal") cfg.weightsUniform = false;
    else throw std::runtime_error("Invalid initial weights distribution type. Possible values: normal, uniform.");
    cfg.weightsUniformMin = (real_t)opt.num("weights_uniform_min"); cfg.weightsUniformMax = (real_t)opt.num("weights_uniform_max");
    cfg.weightsNormalMean = (real_t)opt.num("weights_normal_mean"); cfg.weightsNormalSigma = (real_t)opt.num("weights_normal_sigma");

    const int parallelSequences = (int)opt.num("parallel_sequences");
    const int truncSeq = (int)opt.num("truncate_seq");

    const std::string networkFile = opt.str("continue").empty() ? opt.str("network") : opt.str("continue");      // main.cpp:102
    if (chief) { std::printf("Reading network from '%s'... ", networkFile.c_str()); std::fflush(stdout); }
    std::ifstream nf(networkFile.c_str(), std::ios::binary);
    if (!nf) throw std::runtime_error("Cannot open file");
    std::stringstream nss; nss << nf.rdbuf();
    helpers::JsonDocument netDoc = helpers::parseJson(nss.str());
    if (chief) std::printf("done.\n\n");

    auto load = [&](const char *what, const char *filesKey, const char *fracKey, bool train) -> std::unique_ptr<data_sets::DataSet> {
        const std::vector<std::string> files = opt.list(filesKey);
        if (files.empty()) return nullptr;
        if (chief) { std::printf("Loading %s set '", what); for (size_t i = 0; i < files.size(); ++i) std::printf("%s%s", i ? "' '" : "", files[i].c_str()); std::printf("' ...\n"); std::fflush(stdout); }
        // only the training set is sorted / truncated / shuffled / noised (main.cpp:585-640)
        std::unique_ptr<data_sets::DataSet> ds = data_sets::loadNetCdfDataSet(ctx.p, files, parallelSequences, fracKey ? (real_t)opt.num(fracKey) : 1,
                                                                              train ? truncSeq : 0, train, rank, world);
        // context windows and the target lag apply to every set (the reference reads them from the Configuration singleton)
        ds->setContext((int)opt.num("input_left_context"), (int)opt.num("input_right_context"), (int)opt.num("output_time_lag"));
        if (train) ds->setShuffling(opt.flag("shuffle_fractions"), opt.flag("shuffle_sequences"), blob.seed);
        if (train || !training)          // the reference also adds the input noise to the feed forward input set (main.cpp:610-614)
            ds->setInputNoise((real_t)opt.num("input_noise_sigma"), blob.seed);
        if (chief) {
            std::printf("done.\nLoaded fraction:  %d%%\nSequences:        %d\nSequence lengths: %d..%d\nTotal timesteps:  %d\n\n",
                        (int)(100 * (fracKey ? opt.num(fracKey) : 1)), ds->totalSequences(), ds->minSeqLength(), ds->maxSeqLength(), ds->totalTimesteps());
        }
        return ds;
    };

    std::unique_ptr<data_sets::DataSet> trainingSet, validationSet, testSet, feedForwardSet;
    if (training) {
        trainingSet = load("training", "train_file", "train_fraction", true);
        if (!trainingSet) throw std::runtime_error("No training file given");
        validationSet = load("validation", "val_file", "val_fraction", false);
        testSet = load("test", "test_file", "test_fraction", false);
    } else {
        feedForwardSet = load("feed forward input", "ff_input_file", nullptr, false);
        if (!feedForwardSet) throw std::runtime_error("No feedforward input file given");
    }

    int maxSeqLength = 0;
    for (data_sets::DataSet *d : {trainingSet.get(), validationSet.get(), testSet.get(), feedForwardSet.get()})
        if (d) maxSeqLength = std::max(maxSeqLength, d->maxSeqLength());
    const data_sets::DataSet *shapeSet = training ? trainingSet.get() : feedForwardSet.get();

    if (chief) { std::printf("Creating the neural network... "); std::fflush(stdout); }
    // the input layer takes the spliced pattern size (the reference passes the unspliced size here, main.cpp:148-152, which makes
    // its own context options fail in InputLayer::loadSequences; the spliced size is what the fractions carry)
    NeuralNetwork neuralNetwork(ctx.p, netDoc, parallelSequences, maxSeqLength, shapeSet->fractionInputPatternSize(), shapeSet->outputPatternSize());
    for (data_sets::DataSet *d : {trainingSet.get(), validationSet.get(), testSet.get()})
        if (d && !d->empty() && d->outputPatternSize() != neuralNetwork.postOutputLayer().size())
            throw std::runtime_error("Post output layer size != target pattern size of the data set");
    if (comm) neuralNetwork.setCommunicator(comm);
    if (chief) {
        std::printf("done.\nLayers:\n");
        int i = 0;
        for (const auto &layer : neuralNetwork.layers()) {                                   // main.cpp:646-665
            std::printf("(%d) %s ", i++, layer->type().c_str());
            std::printf("[size: %d", layer->size());
            const layers::TrainableLayer *tl = dynamic_cast<const layers::TrainableLayer *>(layer.get());
            if (tl) std::printf(", bias: %.1lf, weights: %d", (double)tl->bias(), (int)tl->weights().size());
            std::printf("]\n");
        }
        std::printf("Total weights: %d\n\n\n", [&] { int n = 0; for (const auto &l : neuralNetwork.layers()) { auto *tl = dynamic_cast<const layers::TrainableLayer *>(l.get()); if (tl) n += (int)tl->weights().size(); } return n; }());
    }
    const bool classificationTask = dynamic_cast<layers::MulticlassClassificationLayer *>(&neuralNetwork.postOutputLayer()) != nullptr
                                 || dynamic_cast<layers::BinaryClassificationLayer *>(&neuralNetwork.postOutputLayer()) != nullptr;

    if (training) {
        const bool stochastic = opt.flag("stochastic") || opt.flag("hybrid_online_batch");
        optimizers::SteepestDescentOptimizer optimizer(neuralNetwork, (real_t)opt.num("learning_rate"), (real_t)opt.num("momentum"), stochastic);
        const int maxEpochs = (int)opt.num("max_epochs"), maxEpochsNoBest = (int)opt.num("max_epochs_no_best");
        const int validateEvery = (int)opt.num("validate_every"), testEvery = (int)opt.num("test_every");
        optimizer.setWeightNoise((real_t)opt.num("weight_noise_sigma"), blob.seed);
        optimizer.setDataSets(trainingSet.get(), validationSet.get(), testSet.get(), maxEpochs, maxEpochsNoBest, validateEvery, testEvery);
        if (chief) {
            std::printf("Creating the optimizer... done.\nOptimizer type: Steepest descent with momentum\n");
            if (maxEpochs >= 0) std::printf("Max training epochs:       %d\n", maxEpochs);
            std::printf("Max epochs until new best: %d\nValidation error every:    %d\nTest error every:          %d\nLearning rate:             %g\nMomentum:                  %g\n",
                        maxEpochsNoBest, validateEvery, testEvery, opt.num("learning_rate"), opt.num("momentum"));
            if (world > 1) std::printf("Data parallel:             %d processes x %d parallel sequences\n", world, parallelSequences);
            std::printf("\nStarting training...\n\n");
            std::printf(" Epoch | Duration |  Training error  | Validation error |    Test error    | New best \n");
            std::printf("-------+----------+------------------+------------------+------------------+----------\n");
        }
        const bool haveVal = validationSet && !validationSet->empty(), haveTest = testSet && !testSet->empty();
        std::string prefix = opt.str("autosave_prefix");
        if (prefix.empty()) { const std::string &n = opt.str("network"); size_t pos = n.find_last_of('.'); prefix = (pos != std::string::npos && pos > 0) ? n.substr(0, pos) : n; }
        std::string infoRows;
        if (!opt.str("continue").empty()) {                                                    // main.cpp:198-204, 743-757
            if (chief) { std::printf("Restoring state from '%s'... ", opt.str("continue").c_str()); std::fflush(stdout); }
            if (!netDoc.HasMember("info_rows")) throw std::runtime_error("Missing value 'info_rows'");
            infoRows = replaceAll(netDoc["info_rows"].GetString(), ";;;", "\n");
            optimizer.importState(netDoc);
            if (chief) { std::printf("done.\n\n"); std::fputs(infoRows.c_str(), stdout); }
        }
        bool finished = optimizer.finished();
        while (!finished) {
            const char *errFormat = classificationTask ? "%6.2lf%%%10.3lf |" : "%17.3lf |";
            const char *errSpace = "                  |";
            infoRows += printfRow(chief, " %5d | ", optimizer.currentEpoch() + 1);
            const auto t0 = std::chrono::steady_clock::now();
            finished = optimizer.train();
            const double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            infoRows += printfRow(chief, "%8.1lf |", duration);
            auto cell = [&](bool due, double classErr, double err) {
                if (!due) infoRows += printfRow(chief, "%s", errSpace);
                else if (classificationTask) infoRows += printfRow(chief, errFormat, classErr * 100.0, err);
                else infoRows += printfRow(chief, errFormat, err);
            };
            cell(true, optimizer.curTrainingClassError(), optimizer.curTrainingError());
            const bool valDue = haveVal && optimizer.currentEpoch() % validateEvery == 0;
            cell(valDue, optimizer.curValidationClassError(), optimizer.curValidationError());
            cell(haveTest && optimizer.currentEpoch() % testEvery == 0, optimizer.curTestClassError(), optimizer.curTestError());
            if (valDue) {
                if (optimizer.epochsSinceLowestValidationError() == 0) {
                    infoRows += printfRow(chief, "  yes   \n");
                    if (chief && opt.flag("autosave_best") && !finished) saveNetwork(neuralNetwork, prefix + ".best.jsn");
                } else infoRows += printfRow(chief, "  no    \n");
            } else infoRows += printfRow(chief, "        \n");
            if (opt.flag("autosave")) {                                                        // main.cpp:275-277 (state is identical on every rank)
                if (chief) saveState(neuralNetwork, optimizer, opt, infoRows);
            }
        }
        if (chief) {
            std::printf("\n");
            if (optimizer.epochsSinceLowestValidationError() == maxEpochsNoBest) std::printf("No new lowest error since %d epochs. Training stopped.\n", maxEpochsNoBest);
            else std::printf("Maximum number of training epochs reached. Training stopped.\n");
            if (haveVal) std::printf("Lowest validation error: %lf\n", (double)optimizer.lowestValidationError());
            else std::printf("Final training set error: %lf\n", (double)optimizer.curTrainingError());
            std::printf("\nStoring the trained network in '%s'... ", opt.str("save_network").c_str());
            saveNetwork(neuralNetwork, opt.str("save_network"));
            std::printf("done.\n");
        }
    } else {
        if (world > 1) throw std::runtime_error("forward-pass mode runs in a single process");
        const std::vector<real_t> &means = feedForwardSet->outputMeans(), &stdevs = feedForwardSet->outputStdevs();
        const bool unstandardize = opt.flag("revert_std");
        if (unstandardize) std::printf("Outputs will be scaled by mean and standard deviation specified in NC file.\n");
        const int lag = (int)opt.num("output_time_lag");
        const std::string format = opt.str("ff_output_format"), outName = opt.str("ff_output_file");
        if (format != "single_csv" && format != "csv" && format != "htk") throw std::runtime_error("Invalid feedforward format string. Possible values: single_csv, csv, htk.");
        auto value = [&](const std::vector<std::vector<real_t>> &seq, int t, int o) {                 // main.cpp:345-353
            const int T = (int)seq.size();
            real_t v = (t < T - lag) ? seq[t + lag][o] : seq[T - 1][o];
            if (unstandardize) { v *= stdevs[o]; v += means[o]; }
            return v;
        };
        std::ofstream single;
        if (format == "single_csv") { single.open(outName.c_str()); if (!single) throw std::runtime_error("Cannot open '" + outName + "' for writing"); }
        int fracIdx = 0;
        std::shared_ptr<data_sets::DataSetFraction> frac;
        while ((frac = feedForwardSet->getNextFraction())) {
            std::printf("Computing outputs for data fraction %d...", ++fracIdx); std::fflush(stdout);
            neuralNetwork.loadSequences(*frac);
            neuralNetwork.computeForwardPass();
            const std::vector<std::vector<std::vector<real_t>>> outputs = neuralNetwork.getOutputs();
            for (int ps = 0; ps < (int)outputs.size(); ++ps) {
                const std::string &tag = frac->seqInfo(ps).seqTag;
                const int T = (int)outputs[ps].size();
                if (format == "single_csv") {
                    single << tag;
                    for (int t = 0; t < T; ++t) for (int o = 0; o < (int)outputs[ps][t].size(); ++o) single << ';' << value(outputs[ps], t, o);
                    single << '\n';
                } else if (format == "csv") {
                    std::string base = baseName(tag); const size_t dot = base.find_last_of('.');
                    if (dot != std::string::npos && dot > 0) base.erase(dot);
                    const std::string dir = outName + (relativeDir(tag).empty() ? "" : "/" + relativeDir(tag));
                    makeDirs(dir);
                    std::ofstream f((dir + "/" + base + ".csv").c_str());
                    for (int t = 0; t < T; ++t) {
                        for (int o = 0; o < (int)outputs[ps][t].size(); ++o) { if (o) f << ';'; f << value(outputs[ps], t, o); }
                        f << '\n';
                    }
                } else if (T > 0) {                                                                      // HTK, main.cpp:430-478
                    const std::string dir = outName + (relativeDir(tag).empty() ? "" : "/" + relativeDir(tag));
                    makeDirs(dir);
                    std::ofstream f((dir + "/" + baseName(tag) + ".htk").c_str(), std::ios::binary);
                    const int nComps = (int)outputs[ps][0].size();
                    put32(f, (uint32_t)T);
                    put32(f, (uint32_t)(opt.num("feature_period") * 1e4));
                    put16(f, (uint16_t)(nComps * sizeof(float)));
                    put16(f, (uint16_t)opt.num("ff_output_kind"));
                    for (int t = 0; t < T; ++t) for (int o = 0; o < nComps; ++o) { float v = value(outputs[ps], t, o); uint32_t u; std::memcpy(&u, &v, 4); put32(f, u); }
                }
            }
            std::printf(" done.\n");
        }
    }
    if (comm) bl_comm_destroy(comm);
    return 0;
}

} // namespace

int main(int argc, const char **argv)
{
    Options opt;
    try {
        opt.parse(argc, argv);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Error while parsing the command line and/or options file: %s\n", e.what());
        return 1;
    }
    if (opt.v.count("help")) { printHelp(); return 0; }
    if (!opt.str("continue").empty()) {                                                 // options come from the autosave file
        try {
            std::ifstream f(opt.str("continue").c_str(), std::ios::binary);
            if (!f) throw std::runtime_error("Cannot open file");
            std::stringstream ss; ss << f.rdbuf();
            const helpers::JsonDocument doc = helpers::parseJson(ss.str());
            if (!doc.HasMember("configuration")) throw std::runtime_error("Missing string 'configuration'");
            opt.restoreFrom(doc["configuration"].GetString());
        } catch (const std::exception &e) {
            std::fprintf(stderr, "Error while restoring configuration from autosave file: %s\n", e.what());
            return 1;
        }
    }
    try {
        if (opt.flag("list_devices")) {
            std::printf("%d devices found\n", bl_device_count());
            return 0;
        }
        return run(opt);
    } catch (const std::exception &e) {
        std::printf("FAILED: %s\n", e.what());                                         // main.cpp:488-491
        return 2;
    }
}
