// tests/cpp/accessor.h -- the test back door, as in reference tests/accessor.h:12-22: a PrivateAccessor
// specialisation-free template that is a friend of AsRigidAsPossibleDeformation and returns its
// (device-resident) edge-weight matrix as ARAP::SparseMatrix.
#ifndef DEFORM_TEST_ACCESSOR_H
#define DEFORM_TEST_ACCESSOR_H

namespace deform {

template <class ARAP>
class PrivateAccessor {
public:
    static typename ARAP::SparseMatrix cotanWeights(const ARAP &a) {
        return a._edgeWeights;
    }
};

}  // namespace deform

#endif
