// oracle/boost_shim -- see matrix.hpp
#pragma once
