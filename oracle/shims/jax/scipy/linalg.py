from scipy.linalg import expm
