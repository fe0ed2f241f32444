// GoICP_b200 -- the reference's command line (jly_main.cpp:54-179) on the B200 engine:
//     GoICP_b200 <target.mol2> <source.mol2> <NdDownsampled> <config.txt> <output.txt> <pair>
// Same positional arguments (parseInput :181-229), same config.txt keys (readConfig :231-270), same files written:
//   cavitiesN/<id>_sim<pair>N.xyz x2 (normalised clouds, :88-93), <output> (:131-141), <output-stem>_rescaled.txt (:155-156).
// The working directory must hold cavitiesN/ and cfpfh/<id>.cfpfh exactly as for the reference (loadPointCloud :272-314
// exits with -1 when the descriptor file is missing, also when cfpfh=0; kept).
// Host side here: file parsing / formatting.  Device side (through include/goicp_dropin.hpp): centring, max-norm,
// scaling, DT build, registration, rescaled translation.
#include <chrono>

#include "goicp_io.hpp"

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

    std::cout << "Optimal Rotation Matrix:" << std::endl << goicp.optR << std::endl;
    std::cout << "Optimal Translation Vector:" << std::endl << goicp.optT << std::endl;
    std::cout << "Finished in " << time << std::endl << std::endl;

    double Ropt[9], topt[3];
    goicp.optR.getData(Ropt); goicp.optT.getData(topt);
    write_output_file(outputF, time, Ropt, topt, goicp.optError, goicp.Nd - goicp.optComp);

    // rescaleCloud (transformation.cpp:403-417)
    const double meanT[3] = {xMeanT, yMeanT, zMeanT}, meanS[3] = {xMean, yMean, zMean};
    double tr[3];
    t.rescaleTranslation(scale, meanT, meanS, Ropt, topt, tr);
    write_rescaled_file(outputF, time, Ropt, tr, goicp.optError);

    delete[] goicp.pModel; delete[] goicp.pData;
    return 0;
}
