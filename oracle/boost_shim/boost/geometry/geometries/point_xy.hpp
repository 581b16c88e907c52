// oracle/boost_shim -- see boost/geometry.hpp
#pragma once
#include <boost/geometry.hpp>
