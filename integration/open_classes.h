/*
 * open_classes.h — force-included (`g++ -include integration/open_classes.h`) in front of
 * gpu_jobscheduler.cpp ONLY.  The drop-in has to read two things the reference keeps private and offers no
 * accessor for: RolloutJob's {state, score} (mcts/include/mcts.h:82-91) and Quoridor_state's board members
 * (examples/Quoridor/Quoridor.h:32-48; TicTacToe_state::board, TicTacToe.h).  Spelling `class` as `struct`
 * after all standard headers are in changes neither layout nor name mangling, so the object links against
 * the reference's own, unmodified translation units.  A maintainer would add the gpu_pack() hook of
 * INTEGRATION.md (B) instead; this header exists so that the swap needs NO edit of the reference at all.
 */
#ifndef MCTSB_OPEN_CLASSES_H
#define MCTSB_OPEN_CLASSES_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <forward_list>
#include <iomanip>
#include <iostream>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <pthread.h>
#include <stdlib.h>
#define class struct
#endif
