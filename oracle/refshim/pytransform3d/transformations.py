import numpy as np
from distance3d_b200._transforms import (  # noqa: F401
    transform_from, random_transform, concat, invert_transform,
    assert_transform, transform_from_exponential_coordinates)


def vectors_to_points(V):
    return np.hstack((V, np.ones((len(V), 1))))


def transform(A2B, PA):
    return np.dot(PA, A2B.T)
