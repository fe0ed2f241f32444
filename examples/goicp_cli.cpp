// GoICP_b200 -- the reference's command line (jly_main.cpp:54-179) on the B200 engine:
//     GoICP_b200 <target.mol2> <source.mol2> <NdDownsampled> <config.txt> <output.txt> <pair>
// Same positional arguments (parseInput :181-229), same config.txt keys (readConfig :231-270), same files written:
//   cavitiesN/<id>_sim<pair>N.xyz x2 (normalised clouds, :88-93), <output> (:131-141), <output-stem>_rescaled.txt (:155-156).
// The working directory must hold cavitiesN/ and cfpfh/<id>.cfpfh exactly as for the reference (loadPointCloud :272-314
// exits with -1 when the descriptor file is missing, also when cfpfh=0; kept).
// Host side here: file parsing / formatting.  Device side (through include/goicp_dropin.hpp): centring, max-norm,
// scaling, DT build, registration, rescaled translation.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "goicp_dropin.hpp"

using std::string;

// colour codes of the `properties` enum (transformation.hpp:36); unknown atom names map to OG (transformation.cpp:46)
static int atom_colour(const string& name) {
    static const std::map<string, int> table = {{"OG", 8204959}, {"N", 30894}, {"O", 15219528}, {"NZ", 15231913}, {"CZ", 4646984},
                                                {"CA", 16741671}, {"DU", 7566712}, {"OD1", 0}, {"C", 1}};
    auto it = table.find(name);
    return it == table.end() ? 8204959 : it->second;
}

// The @<TRIPOS>ATOM block of a mol2 file: id, name, x, y, z per atom (transformation.cpp:282-306 keeps every atom
// whose five leading fields parse and drops the first record that does not).
static std::vector<point4D> read_mol2_atoms(const string& path) {
    std::vector<point4D> cloud;
    std::ifstream in(path);
    string line;
    while (std::getline(in, line))
        if (line.find("@<TRIPOS>ATOM") != string::npos) break;
    string id, name;
    point4D p{};
    while (in >> id >> name >> p.x >> p.y >> p.z) {
        p.c = atom_colour(name);
        cloud.push_back(p);
        std::getline(in, line);   // rest of the atom record
    }
    return cloud;
}

// writeNormalizedMolCloudFile (transformation.cpp:340-350): default ostream formatting = 6 significant digits
static void write_xyzc(const string& path, const std::vector<point4D>& cloud) {
    std::ofstream out(path);
    out << cloud.size() << std::endl;
    for (const point4D& p : cloud) out << p.x << " " << p.y << " " << p.z << " " << p.c << std::endl;
}

// ConfigMap (ConfigMap.cpp:3-151): key=value, '#' comments, tokens split on " =;", lines without exactly two tokens ignored
static std::map<string, string> read_config(const string& path) {
    std::ifstream in(path);
    if (!in.is_open()) { std::cout << "Unable to open config file '" << path << "'" << std::endl; exit(-2); }
    std::map<string, string> m;
    string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const size_t h = line.find('#');
        if (h != string::npos) line = line.substr(0, h);
        std::vector<string> tok; string cur;
        for (char ch : line) { if (ch == ' ' || ch == '=' || ch == ';') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } } else cur += ch; }
        if (!cur.empty()) tok.push_back(cur);
        if (tok.size() == 2) m[tok[0]] = tok[1];
    }
    return m;
}
static double cfgF(const std::map<string, string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0.0 : atof(it->second.c_str()); }
static int cfgI(const std::map<string, string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0 : atoi(it->second.c_str()); }

// loadPointCloud (jly_main.cpp:272-314): the normalised xyzc text + cfpfh/<id>.cfpfh (41 bins per point)
static void load_cloud(const string& fname, int& N, POINT3D** p) {
    const string fp = "cfpfh/" + fname.substr(fname.find("/") + 1, fname.find_last_of("_") - fname.find("/") - 1) + ".cfpfh";
    std::ifstream in(fname), fin(fp);
    if (!in.is_open()) { std::cout << "Unable to open point file '" << fname << "'" << std::endl; exit(-1); }
    in >> N;
    *p = new POINT3D[N];
    if (!fin.is_open()) { std::cout << "Unable to open fpfh file '" << fp << "'" << std::endl; exit(-1); }
    for (int i = 0; i < N; i++) {
        POINT3D& q = (*p)[i];
        in >> q.x >> q.y >> q.z >> q.c;
        q.neighbors = 0; q.density = 0;
        q.cfpfh.resize(41);
        for (int j = 0; j < 41; j++) { float b = 0; fin >> b; q.cfpfh[j] = b; }
    }
}

static string stem_between(const string& s) { return s.substr(s.find("/") + 1, s.find(".") - s.find("/") - 1); }   // jly_main.cpp:68-69

int main(int argc, char** argv) {
    string modelF = "model.txt", dataF = "data.txt", configF = "config.txt", outputF = "output.txt";   // defaults :50-53
    int NdDown = 0, pair = 1;
    if (argc > 6) pair = atoi(argv[6]);
    if (argc > 5) outputF = argv[5];
    if (argc > 4) configF = argv[4];
    if (argc > 3) NdDown = atoi(argv[3]);
    if (argc > 2) dataF = argv[2];
    if (argc > 1) modelF = argv[1];

    GoICP goicp;
    const auto cfg = read_config(configF);
    goicp.MSEThresh = (float)cfgF(cfg, "MSEThresh");
    goicp.initNodeRot.a = (float)cfgF(cfg, "rotMinX"); goicp.initNodeRot.b = (float)cfgF(cfg, "rotMinY"); goicp.initNodeRot.c = (float)cfgF(cfg, "rotMinZ");
    goicp.initNodeRot.w = (float)cfgF(cfg, "rotWidth");
    goicp.initNodeTrans.x = (float)cfgF(cfg, "transMinX"); goicp.initNodeTrans.y = (float)cfgF(cfg, "transMinY"); goicp.initNodeTrans.z = (float)cfgF(cfg, "transMinZ");
    goicp.initNodeTrans.w = (float)cfgF(cfg, "transWidth");
    goicp.trimFraction = (float)cfgF(cfg, "trimFraction");
    goicp.regularization = (float)cfgF(cfg, "regularization");
    goicp.regularizationNeighbors = (float)cfgF(cfg, "regularizationNeighbors");
    goicp.regularizationFPFH = (float)cfgF(cfg, "regularizationFPFH");
    goicp.cfpfh = cfgI(cfg, "cfpfh"); goicp.norm = cfgI(cfg, "norm"); goicp.ponderation = cfgI(cfg, "ponderation");
    if (goicp.trimFraction < 0.001) goicp.doTrim = false;
    goicp.dt.SIZE = cfgI(cfg, "distTransSize"); goicp.dt.expandFactor = cfgF(cfg, "distTransExpandFactor");

    const string proteinData = stem_between(dataF), proteinModel = stem_between(modelF);
    Transformation t;
    std::vector<point4D> cloudSource = read_mol2_atoms(dataF), cloudTarget = read_mol2_atoms(modelF);
    if (cloudSource.empty() || cloudTarget.empty()) { std::cout << "Unable to read atoms from '" << (cloudSource.empty() ? dataF : modelF) << "'" << std::endl; return -1; }
    double xMean, yMean, zMean, xMeanT, yMeanT, zMeanT;
    const double scaleSource = t.normalizeMolCloud(cloudSource, xMean, yMean, zMean);
    const double scaleTarget = t.normalizeMolCloud(cloudTarget, xMeanT, yMeanT, zMeanT);
    const double scale = scaleSource >= scaleTarget ? scaleSource : scaleTarget;   // :85
    t.scaleCloud(cloudSource, scale);
    t.scaleCloud(cloudTarget, scale);
    const string sourceN = "cavitiesN/" + proteinData + "_sim" + std::to_string(pair) + "N.xyz";
    const string targetN = "cavitiesN/" + proteinModel + "_sim" + std::to_string(pair) + "N.xyz";
    write_xyzc(sourceN, cloudSource);
    write_xyzc(targetN, cloudTarget);

    load_cloud(targetN, goicp.Nm, &goicp.pModel);   // the text round trip is part of the reference's numerics
    load_cloud(sourceN, goicp.Nd, &goicp.pData);

    std::cout << "Building Distance Transform..." << std::flush;
    auto c0 = std::chrono::steady_clock::now();
    goicp.BuildDT();
    std::cout << std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count() << "s" << std::endl;
    if (NdDown > 0) goicp.Nd = NdDown;
    std::cout << "Model ID: " << modelF << " (" << goicp.Nm << "), Data ID: " << dataF << " (" << goicp.Nd << ")" << std::endl;
    std::cout << "Registering..." << std::endl;
    c0 = std::chrono::steady_clock::now();
    goicp.Register();
    const double time = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
    std::cout << goicp.Trace();

    auto matrix_rows = [](const double* v, int rows, int cols) {   // Matrix operator<< (matrix.cpp:812-827): "%12.7f " per entry
        std::ostringstream o; char b[64];
        for (int i = 0; i < rows; i++) { for (int j = 0; j < cols; j++) { snprintf(b, sizeof b, "%12.7f ", v[i * cols + j]); o << b; } if (i < rows - 1) o << std::endl; }
        return o.str();
    };
    std::cout << "Optimal Rotation Matrix:" << std::endl << matrix_rows(goicp.optR, 3, 3) << std::endl;
    std::cout << "Optimal Translation Vector:" << std::endl << matrix_rows(goicp.optT, 3, 1) << std::endl;
    std::cout << "Finished in " << time << std::endl << std::endl;

    std::ofstream ofile(outputF.c_str());
    ofile << "Time: " << time << std::endl;
    ofile << "Rotation Matrix: " << std::endl << matrix_rows(goicp.optR, 3, 3) << std::endl;
    ofile << "Translation Vector: " << std::endl << matrix_rows(goicp.optT, 3, 1) << std::endl;
    ofile << "Error: " << goicp.optError << std::endl;
    ofile << "Compatibilities: " << goicp.Nd - goicp.optComp << std::endl;
    ofile.close();

    // rescaleCloud (transformation.cpp:403-417)
    const double meanT[3] = {xMeanT, yMeanT, zMeanT}, meanS[3] = {xMean, yMean, zMean};
    double tr[3];
    t.rescaleTranslation(scale, meanT, meanS, goicp.optR, goicp.optT, tr);
    std::ofstream rf(outputF.substr(0, outputF.find(".")) + "_rescaled.txt");
    rf << "Time: " << time << std::endl;
    rf << "Rotation Matrix:" << std::endl << "   " << goicp.optR[0] << "   " << goicp.optR[1] << "   " << goicp.optR[2] << std::endl;
    rf << "   " << goicp.optR[3] << "   " << goicp.optR[4] << "   " << goicp.optR[5] << std::endl;
    rf << "   " << goicp.optR[6] << "   " << goicp.optR[7] << "   " << goicp.optR[8] << std::endl;
    rf << "Translation Vector:" << std::endl << "   " << tr[0] << std::endl << "   " << tr[1] << std::endl << "   " << tr[2] << std::endl;
    rf << "Error: " << goicp.optError << std::endl;
    rf.close();

    delete[] goicp.pModel; delete[] goicp.pData;
    return 0;
}
