// shim: included by the reference but unused
#pragma once
