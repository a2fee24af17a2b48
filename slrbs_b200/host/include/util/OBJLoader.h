// util/OBJLoader.h -- include/util/OBJLoader.h of the reference: OBJLoader::load(filename, meshV, meshF) reads the 'v x y z'
// lines and the first index of each 'f a/b/c' token (1-based, triangles), src/util/OBJLoader.cpp:34-122. Header-only here.
#pragma once
#include <Eigen/Dense>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

class OBJLoader {
public:
    static bool load(const std::string& filename, Eigen::MatrixXf& meshV, Eigen::MatrixXi& meshF)
    {
        std::FILE* f = std::fopen(filename.c_str(), "r");
        if (!f) { std::fprintf(stderr, "Error: Failed to open file %s for reading!\n", filename.c_str()); return false; }
        std::vector<float> verts;
        std::vector<int> inds;
        char line[1024];
        while (std::fgets(line, sizeof line, f)) {
            if (line[0] == 'v' && line[1] == ' ') {
                float x = 0.f, y = 0.f, z = 0.f;
                std::sscanf(line + 2, "%f %f %f", &x, &y, &z);
                verts.push_back(x); verts.push_back(y); verts.push_back(z);
            } else if (line[0] == 'f' && line[1] == ' ') {
                for (char* tok = std::strtok(line + 2, " \t\r\n"); tok; tok = std::strtok(nullptr, " \t\r\n")) inds.push_back(std::atoi(tok));
            }
        }
        std::fclose(f);
        const int nv = (int)verts.size() / 3, nf = (int)inds.size() / 3;
        meshV.resize(nv, 3);
        for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) meshV(i, k) = verts[3 * i + k];
        meshF.resize(nf, 3);
        for (int i = 0; i < nf; ++i) for (int k = 0; k < 3; ++k) meshF(i, k) = inds[3 * i + k] - 1;
        return true;
    }
};
