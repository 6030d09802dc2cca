from numpy import *
