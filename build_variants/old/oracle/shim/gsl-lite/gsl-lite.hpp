// shim: mesh.h includes gsl-lite but the hot-path TUs use nothing from it.
#pragma once
