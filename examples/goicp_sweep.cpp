// GoICP_b200_sweep -- the reference's dataset driver (bo1_GoICP.py:6-54: one ./GoICP process per TSV row) as ONE in-process
// batch on the B200 engine:
//     GoICP_b200_sweep <pairs.tsv> <config.txt> [prefix=similar] [first_index=1]
// TSV columns 3 and 4 name the source and target cavity (bo1_GoICP.py:14-16: cavities/<id>_cavity6.mol2); row K is pair
// first_index + K - 1.  Per pair the same files as the per-pair command line are written: cavitiesN/<id>_sim<K>N.xyz x2
// (jly_main.cpp:88-93), output/<prefix><K>.txt (:131-141) and output/<prefix><K>_rescaled.txt (:155-156).  All pairs are
// registered by one goicp_register_batch call (include/goicp_b200.h): DT builds, Initialize and the searches of every
// pair share the GPU instead of running one after the other.
#include <chrono>

#include "goicp_io.hpp"

struct PairFiles {
    string source, target, outputF;
    int pair;
    double scale, meanS[3], meanT[3];
    int Nm = 0, Nd = 0;
    POINT3D *pModel = nullptr, *pData = nullptr;
    std::vector<float> mxyz, dxyz, mf, df;
    std::vector<int32_t> mc, dc;
};

static void flatten(const POINT3D* p, int n, std::vector<float>& xyz, std::vector<int32_t>& c, std::vector<float>& f) {
    xyz.resize(3 * (size_t)n); c.resize(n); f.resize(41 * (size_t)n);
    for (int i = 0; i < n; i++) {
        xyz[3 * i] = p[i].x; xyz[3 * i + 1] = p[i].y; xyz[3 * i + 2] = p[i].z; c[i] = p[i].c;
        for (int k = 0; k < 41; k++) f[41 * (size_t)i + k] = p[i].cfpfh[k];
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { std::cout << "usage: GoICP_b200_sweep <pairs.tsv> <config.txt> [prefix=similar] [first_index=1]" << std::endl; return 1; }
    const string tsvF = argv[1], configF = argv[2], prefix = argc > 3 ? argv[3] : "similar";
    const int first = argc > 4 ? atoi(argv[4]) : 1;
    const goicp_params params = params_from_config(read_config(configF));

    std::ifstream tsv(tsvF);
    if (!tsv.is_open()) { std::cout << "Unable to open pair list '" << tsvF << "'" << std::endl; return -1; }
    std::vector<PairFiles> pairs;
    string line;
    while (std::getline(tsv, line)) {   // bo1_GoICP.py:10-18: stop at the first blank line
        std::istringstream ls(line);
        std::vector<string> col; string w;
        while (ls >> w) col.push_back(w);
        if (col.empty()) break;
        if (col.size() < 4) { std::cout << "pair list row " << pairs.size() + 1 << " has fewer than 4 columns" << std::endl; return -1; }
        PairFiles pf;
        pf.source = "cavities/" + col[2] + "_cavity6.mol2"; pf.target = "cavities/" + col[3] + "_cavity6.mol2";
        pf.pair = first + (int)pairs.size();
        pf.outputF = "output/" + prefix + std::to_string(pf.pair) + ".txt";
        pairs.push_back(pf);
    }
    if (pairs.empty()) { std::cout << "no pairs in '" << tsvF << "'" << std::endl; return -1; }

    // host side of jly_main.cpp:61-104 per pair: mol2 -> centred, jointly scaled clouds -> text round trip + descriptors
    Transformation t;
    for (PairFiles& pf : pairs) {
        std::vector<point4D> cloudSource = read_mol2_atoms(pf.source), cloudTarget = read_mol2_atoms(pf.target);
        if (cloudSource.empty() || cloudTarget.empty()) { std::cout << "Unable to read atoms from '" << (cloudSource.empty() ? pf.source : pf.target) << "'" << std::endl; return -1; }
        const double sS = t.normalizeMolCloud(cloudSource, pf.meanS[0], pf.meanS[1], pf.meanS[2]);
        const double sT = t.normalizeMolCloud(cloudTarget, pf.meanT[0], pf.meanT[1], pf.meanT[2]);
        pf.scale = sS >= sT ? sS : sT;   // :85
        t.scaleCloud(cloudSource, pf.scale); t.scaleCloud(cloudTarget, pf.scale);
        const string sourceN = "cavitiesN/" + stem_between(pf.source) + "_sim" + std::to_string(pf.pair) + "N.xyz";
        const string targetN = "cavitiesN/" + stem_between(pf.target) + "_sim" + std::to_string(pf.pair) + "N.xyz";
        write_xyzc(sourceN, cloudSource); write_xyzc(targetN, cloudTarget);
        load_cloud(targetN, pf.Nm, &pf.pModel); load_cloud(sourceN, pf.Nd, &pf.pData);
        flatten(pf.pModel, pf.Nm, pf.mxyz, pf.mc, pf.mf); flatten(pf.pData, pf.Nd, pf.dxyz, pf.dc, pf.df);
        delete[] pf.pModel; delete[] pf.pData;
    }

    std::vector<goicp_pair_desc> desc(pairs.size());
    for (size_t k = 0; k < pairs.size(); k++) {
        PairFiles& pf = pairs[k];
        desc[k].model_xyz = pf.mxyz.data(); desc[k].model_c = pf.mc.data(); desc[k].model_fpfh = pf.mf.data(); desc[k].Nm = pf.Nm;
        desc[k].data_xyz = pf.dxyz.data(); desc[k].data_c = pf.dc.data(); desc[k].data_fpfh = pf.df.data(); desc[k].NdAll = pf.Nd;
        desc[k].Nd = pf.Nd;   // bo1_GoICP.py:46-51 passes the source's atom count as NdDownsampled
    }
    goicp_handle h = nullptr;
    if (goicp_create(&h, 0, nullptr) != GOICP_OK) { std::cout << "goicp_create: " << goicp_last_error(nullptr) << std::endl; return -3; }
    std::vector<goicp_result> res(pairs.size());
    std::cout << "Registering " << pairs.size() << " pairs..." << std::endl;
    const auto c0 = std::chrono::steady_clock::now();
    if (goicp_register_batch(h, &params, (int32_t)pairs.size(), desc.data(), res.data()) != GOICP_OK) { std::cout << "goicp_register_batch: " << goicp_last_error(h) << std::endl; return -3; }
    const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();

    for (size_t k = 0; k < pairs.size(); k++) {
        const PairFiles& pf = pairs[k]; const goicp_result& r = res[k];
        const double time = r.seconds_register;
        write_output_file(pf.outputF, time, r.R, r.t, r.optError, pf.Nd - r.optComp);
        double tr[3];
        t.rescaleTranslation(pf.scale, pf.meanT, pf.meanS, r.R, r.t, tr);
        write_rescaled_file(pf.outputF, time, r.R, tr, r.optError);
        std::cout << prefix << pf.pair << ": " << pf.target << " <- " << pf.source << "  Error: " << r.optError << "  Compatibilities: " << pf.Nd - r.optComp << std::endl;
        // protein-level post-processing (the block of jly_main.cpp:159-172, README.md:25): when chains/<source>_protein.mol2 is there, the
        // rescaled transform is applied to the whole protein (rot/rot_<protein>) and its backbone RMSD against
        // ref_proteins/<source>.<target>/aligned_<protein> is appended to resultsRMSD.txt.  Like the reference, the directories
        // cavitiesR/ and rot/ must exist.
        const string src6 = stem_between(pf.source).substr(0, 6), tgt6 = stem_between(pf.target).substr(0, 6);
        const string protein = src6 + "_protein.mol2";
        if (std::ifstream("chains/" + protein).good()) {
            {   // applyTransformationProtein reads cavitiesR/similar<pair>.txt = the rescaled result (transformation.cpp:470)
                std::ifstream a(pf.outputF.substr(0, pf.outputF.find(".")) + "_rescaled.txt", std::ios::binary);
                std::ofstream b("cavitiesR/similar" + std::to_string(pf.pair) + ".txt", std::ios::binary);
                b << a.rdbuf();
            }
            std::ofstream rotOut("rot/rot_" + protein);
            t.applyTransformationProtein(rotOut, "chains/" + protein, pf.pair);
            std::ifstream aligned("ref_proteins/" + src6 + "." + tgt6 + "/aligned_" + protein), rot("rot/rot_" + protein);
            if (aligned.is_open() && rot.is_open()) {
                const float rmsd = t.computeRMSD(aligned, rot);
                std::ofstream rf("resultsRMSD.txt", std::ios::app);
                if (rf.is_open()) rf << pf.pair << "\t" << src6 << "\t" << tgt6 << "\t" << std::to_string(rmsd) << std::endl;
                std::cout << "---> RMSD: " << rmsd << std::endl;
            }
        }
    }
    std::cout << "Finished " << pairs.size() << " pairs in " << total << " s (" << pairs.size() / total << " pairs/s)" << std::endl;
    goicp_destroy(h);
    return 0;
}
