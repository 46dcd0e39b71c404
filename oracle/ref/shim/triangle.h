// TEST INFRASTRUCTURE (oracle build only).
// Declarations of Shewchuk's Triangle API so that the reference's Mesh.cpp
// compiles; only referenced under `#if DIM == 2`, never linked for DIM == 3.
#pragma once
#ifndef REAL
#define REAL double
#endif
#ifndef VOID
#define VOID void
#endif
extern "C" {
struct triangulateio {
    REAL* pointlist; REAL* pointattributelist; int* pointmarkerlist;
    int numberofpoints; int numberofpointattributes;
    int* trianglelist; REAL* triangleattributelist; REAL* trianglearealist;
    int* neighborlist; int numberoftriangles; int numberofcorners;
    int numberoftriangleattributes;
    int* segmentlist; int* segmentmarkerlist; int numberofsegments;
    REAL* holelist; int numberofholes;
    REAL* regionlist; int numberofregions;
    int* edgelist; int* edgemarkerlist; REAL* normlist; int numberofedges;
};
void triangulate(char*, struct triangulateio*, struct triangulateio*, struct triangulateio*);
void trifree(VOID* memptr);
}
