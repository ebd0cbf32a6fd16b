// stand-in for the reference's Model.h: the public interface of class Model (Model.h:108-163), declarations only
#pragma once
#include <Eigen/Dense>
#include <opencv2/core/mat.hpp>
class Model {
   public:
    Model(int x, int y, int z, float size);
    int getX();
    int getY();
    int getZ();
    float getSize();
    Vector4f get(int x, int y, int z);
    void set(int x, int y, int z, Vector4f v);
    void see(int x, int y, int z);
    bool isInner(int x, int y, int z);
    void handleUnseen();
};
