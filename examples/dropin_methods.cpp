// A caller of the methods north_star names on the drop-in surface: GoICP::Initialize / OuterBnB / InnerBnB / ICP / Clear,
// DT3D::emptyCells / cellPoints, Matrix (include/goicp_dropin.hpp).  Prints key=value lines that tests/test_gpu_parity.py compares with
// the python binding's results for the same calls.
//     dropin_methods <model: N, then "x y z c f1..f41" rows> <data: same> <Nd>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include "goicp_dropin.hpp"

static POINT3D* load(const char* path, int& n) {
    std::ifstream in(path);
    in >> n;
    POINT3D* p = new POINT3D[n];
    for (int i = 0; i < n; i++) { in >> p[i].x >> p[i].y >> p[i].z >> p[i].c; p[i].cfpfh.resize(41); for (int k = 0; k < 41; k++) in >> p[i].cfpfh[k]; p[i].neighbors = 0; p[i].density = 0; }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    GoICP goicp;
    goicp.pModel = load(argv[1], goicp.Nm);
    goicp.pData = load(argv[2], goicp.Nd);
    // shipped config.txt
    goicp.MSEThresh = 0.01f;
    goicp.initNodeRot.a = goicp.initNodeRot.b = goicp.initNodeRot.c = -3.1416f; goicp.initNodeRot.w = 6.2832f;
    goicp.initNodeTrans.x = goicp.initNodeTrans.y = goicp.initNodeTrans.z = -0.5f; goicp.initNodeTrans.w = 1.0f;
    goicp.trimFraction = 0; goicp.doTrim = false;
    goicp.regularization = 0.0005f; goicp.norm = 2; goicp.ponderation = 1;
    goicp.dt.SIZE = 20; goicp.dt.expandFactor = 2.0;
    goicp.BuildDT();
    goicp.Nd = atoi(argv[3]);
    goicp.Register();
    printf("register_optError=%.9g\nregister_optComp=%d\n", goicp.optError, goicp.optComp);
    { std::ostringstream o; o << goicp.optR; std::string rows = o.str(); printf("matrix_row0=%s\n", rows.substr(0, rows.find('\n')).c_str()); }

    goicp.Initialize();
    printf("SSEThresh=%.9g\ninlierNum=%d\nmaxRotDis_3_5=%.9g\nweight_7=%.9g\n", goicp.SSEThresh, goicp.inlierNum, goicp.maxRotDis[3][5], goicp.weights[7]);
    goicp.optError = 30.0f;
    TRANSNODE tn{};
    printf("inner_ub=%.9g\n", goicp.InnerBnB(NULL, &tn));
    printf("inner_lb=%.9g\n", goicp.InnerBnB(goicp.maxRotDis[2], NULL));
    Matrix R = Matrix::eye(3), t(3, 1);
    printf("icp_err=%.9g\n", goicp.ICP(R, t));
    printf("icp_R00=%.9g\n", R.val[0][0]);
    printf("outer_optError=%.9g\n", goicp.OuterBnB());
    goicp.Clear();

    const EMPTYCELL e = goicp.dt.emptyCells[3][4][5];
    const CELL& cell = goicp.dt.cellPoints[e.z][e.y][e.x];
    printf("empty_x=%d\nempty_y=%d\nempty_z=%d\ncell_c=%d\ncell_npoints=%d\n", e.x, e.y, e.z, cell.c, (int)cell.points.size());
    delete[] goicp.pModel; delete[] goicp.pData;
    return 0;
}
