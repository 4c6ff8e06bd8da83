// shim: empty (reference includes it, uses nothing from it on the hot path)
#pragma once
