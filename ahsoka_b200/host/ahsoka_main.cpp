// ahsoka_main.cpp — `Ahsoka phase` / `Ahsoka only-bubbles` with the phasing step behind the
// C ABI.  Same options, same files, same stdout banners as reference src/polyassembly.cpp:22-176;
// the reference's own translation units (graph.cpp, alignmentreader.cpp, argumentparser.cpp,
// chainstoreadset.cpp) are linked unchanged for parsing, bubble/chain detection and allele-path
// enumeration, exactly as BASELINE.json's north_star prescribes.  `-t N` with N > 1 runs the
// same single batch call (the reference's two-thread experiment, polyassembly.cpp:190-222,
// only ever processed the ten largest chains and is not a parity target).
#include <algorithm>
#include <fstream>
#include <functional>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "alignmentreader.hpp"
#include "argumentparser.hpp"
#include "graph.hpp"

using std::cerr; using std::cout; using std::endl; using std::string;
typedef std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>> ChainAlleles;

ChainAlleles ChainsToReadsetDetailed(Graph graph);     // reference src/chainstoreadset.cpp:161
void alignmentsToReadset(AlignmentReader&, Graph&, ChainAlleles&, string, bool, std::vector<std::pair<int, int>>&, std::mutex&);

int main(int argc, char* argv[]) {
    cerr << "Ahsoka: Haplotype assembly for diploid and polyploid genomes based on HiFi and ultra-long ONT data" << endl;
    cerr << "author: Rebecca Serra Mari" << endl;
    ArgumentParser argparser;
    argparser.add_command("Ahsoka [options] -g <graph.gfa> -a <alignments.gaf>");
    argparser.add_subcommand("only-bubbles", {'g', 'o'}, {'t'});
    argparser.add_subcommand("phase", {'g', 'a', 'o'}, {'s', 't'});
    string cmd = argparser.get_subcommand(argc, argv);
    if (cmd == "phase") {
        argparser.add_mandatory_argument('g', "genome assembly graph in gfa format, e.g. by hifiasm");
        argparser.add_mandatory_argument('a', "alignments of ONT reads to the assembly graph, in gaf format");
        argparser.add_mandatory_argument('o', "output folder to store output files");
        argparser.add_optional_argument('s', "", "additional long range phasing information (StrandSeq)");
        argparser.add_optional_argument('t', "1", "number of threads to use");
    } else if (cmd == "only-bubbles") {
        argparser.add_mandatory_argument('g', "genome assembly graph in gfa format, e.g. by hifiasm");
        argparser.add_mandatory_argument('o', "output folder");
        argparser.add_optional_argument('t', "1", "number of threads to use");
    }
    argparser.add_optional_argument('k', "", "kmerfile to be written to during the kmer counting");
    argparser.add_optional_argument('c', "", "outfile for summed up unique kmer counts in short read sample");
    try { argparser.parse(argc, argv); }
    catch (const std::runtime_error& e) { argparser.print_help(); cerr << e.what() << endl; return 1; }
    catch (const std::exception& e) { return 0; }

    const string gfafile = argparser.get_arg_parameter('g');
    const string alignmentfile = argparser.get_arg_parameter('a');
    const string prefix = argparser.get_arg_parameter('o');
    const int threads = std::stoi(argparser.get_arg_parameter('t'));
    Graph graph = Graph::ReadGraph(gfafile);
    cout << "number of threads used: " << threads << endl;
    cout << "threads available: " << std::thread::hardware_concurrency() << endl;
    cout << "Step 1: Graph with " << graph.nodes.size() << " nodes read" << endl;
    graph.findBubbles();
    cout << "Step 2: Bubbles read" << endl;
    cout << "Number of bubble chains: " << graph.chains.size() << endl;
    {
        std::ofstream bubblefile(prefix + "-bubbleinfo.txt");
        for (auto& chain : graph.chains) {     // graph.chains[i].id == i (graph.cpp:351-365): no getChain() scan needed
            bubblefile << "chain id: " << chain.id << "size: " << chain.bubbles.size() << endl;
            for (auto& bubble : chain.bubbles) {
                bubblefile << "bubble id: " << bubble.id << endl << "node id: ";
                for (auto& node : bubble.getNodes()) bubblefile << node.node_id << ",";
                bubblefile << endl;
            }
        }
    }
    if (cmd == "only-bubbles") return 0;

    AlignmentReader alignmentreader;
    alignmentreader.readAlignmentfile(alignmentfile, graph);
    cout << "Step 3: Alignments read" << endl;
    cout << "Number of alignments: " << alignmentreader.alignments.size() << endl;
    ChainAlleles chainpathToAlleles = ChainsToReadsetDetailed(graph);
    cout << "Step 4: Chain paths computed " << endl;
    cout << "Number of chain paths: " << chainpathToAlleles.size() << endl;
    std::vector<std::pair<int, int>> size_sorting;
    for (auto& chainmap : chainpathToAlleles) size_sorting.emplace_back(chainmap.second.size(), chainmap.first);
    std::sort(size_sorting.begin(), size_sorting.end(), std::greater<>());
    cout << "Step 5: Phasing processed" << endl;
    cout << "single thread" << endl;
    cout << "size sorting: " << size_sorting.size() << endl;
    std::mutex g_display_mutex;
    alignmentsToReadset(alignmentreader, graph, chainpathToAlleles, prefix, false, size_sorting, g_display_mutex);
    return 0;
}
