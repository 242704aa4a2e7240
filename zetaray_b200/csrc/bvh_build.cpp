// bvh_build.cpp -- host builder of the 8-wide quantised BVH (see zr_bvh.h). Plain C++: no CUDA calls.
#include "zr_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace zr
{
namespace
{
    // ------------------------------------------------------------------------------------------
    // host BVH builder
    // ------------------------------------------------------------------------------------------
    struct AABB
    {
        float lo[3], hi[3];
        void reset() { for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
        void grow(const AABB& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
        void grow(const float p[3]) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
        float area() const
        {
            float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
            if (dx < 0) return 0;
            return 2.0f * (dx * dy + dy * dz + dz * dx);
        }
    };

    struct BNode { AABB box; int left = -1, right = -1; uint32_t first = 0, count = 0; };

    struct Builder
    {
        std::vector<AABB> triBox;
        std::vector<float> centroid;    // 3 per tri
        std::vector<uint32_t> order;
        std::vector<BNode> nodes;

        int build(uint32_t first, uint32_t count)
        {
            BNode n;
            n.box.reset();
            AABB cb; cb.reset();
            for (uint32_t i = first; i < first + count; i++)
            {
                n.box.grow(triBox[order[i]]);
                cb.grow(&centroid[order[i] * 3]);
            }
            n.first = first; n.count = count;
            const int idx = (int)nodes.size();
            nodes.push_back(n);
            if (count <= 3)
                return idx;
            // binned SAH over the widest centroid axis (try all 3)
            const int NB = 16;
            float bestCost = INFINITY; int bestAxis = -1; int bestSplit = -1;
            for (int a = 0; a < 3; a++)
            {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0)) continue;
                AABB bb[NB]; uint32_t bc[NB];
                for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
                for (uint32_t i = first; i < first + count; i++)
                {
                    int b = (int)((centroid[order[i] * 3 + a] - cb.lo[a]) / ext * NB);
                    b = std::min(std::max(b, 0), NB - 1);
                    bb[b].grow(triBox[order[i]]); bc[b]++;
                }
                AABB r; r.reset();
                float rArea[NB]; uint32_t rCnt[NB]; uint32_t c = 0;
                for (int b = NB - 1; b > 0; b--) { r.grow(bb[b]); c += bc[b]; rArea[b] = r.area(); rCnt[b] = c; }
                AABB l; l.reset(); c = 0;
                for (int b = 0; b < NB - 1; b++)
                {
                    l.grow(bb[b]); c += bc[b];
                    if (c == 0 || rCnt[b + 1] == 0) continue;
                    const float cost = l.area() * (float)c + rArea[b + 1] * (float)rCnt[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
                }
            }
            uint32_t mid;
            if (bestAxis < 0)
                mid = first + count / 2;        // all centroids coincide: split by index
            else
            {
                const float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
                auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                    int b = (int)((centroid[t * 3 + bestAxis] - cb.lo[bestAxis]) / ext * NB);
                    b = std::min(std::max(b, 0), NB - 1);
                    return b <= bestSplit;
                });
                mid = (uint32_t)(it - order.begin());
                if (mid == first || mid == first + count)
                    mid = first + count / 2;
            }
            const int l = build(first, mid - first);
            const int r = build(mid, first + count - mid);
            nodes[idx].left = l; nodes[idx].right = r;
            return idx;
        }
    };

    // stackBelow: entries on the traversal stack while this node is being processed (unvisited siblings of its ancestors,
    // worst case = this node popped first among its siblings)
    void emit_wide(const Builder& b, int binIdx, uint32_t outIdx, BvhBuild& out, uint32_t depth, uint32_t stackBelow)
    {
        out.maxDepth = std::max(out.maxDepth, depth);
        // gather up to 8 children by repeatedly opening the child with the largest area
        std::vector<int> kids;
        const BNode& root = b.nodes[binIdx];
        if (root.left < 0) kids.push_back(binIdx);
        else { kids.push_back(root.left); kids.push_back(root.right); }
        while (kids.size() < 8)
        {
            int bestK = -1; float bestA = -1;
            for (size_t k = 0; k < kids.size(); k++)
            {
                const BNode& c = b.nodes[kids[k]];
                if (c.left < 0) continue;
                const float a = c.box.area();
                if (a > bestA) { bestA = a; bestK = (int)k; }
            }
            if (bestK < 0) break;
            const BNode c = b.nodes[kids[bestK]];
            kids[bestK] = c.left;
            kids.push_back(c.right);
        }
        BVH8Node n;
        memset(&n, 0, sizeof(n));
        const AABB& box = root.box;
        n.px = box.lo[0]; n.py = box.lo[1]; n.pz = box.lo[2];
        uint8_t* ex[3] = { &n.ex, &n.ey, &n.ez };
        float scale[3];
        for (int a = 0; a < 3; a++)
        {
            const float ext = std::max(box.hi[a] - box.lo[a], 1e-30f);
            int e = (int)std::ceil(std::log2(ext / 255.0f));
            // make sure 255 * 2^e covers the extent even after rounding
            while (std::ldexp(255.0f, e) < ext) e++;
            e = std::min(std::max(e, -126), 127);
            *ex[a] = (uint8_t)(e + 127);
            scale[a] = std::ldexp(1.0f, e);
        }
        // internal children first get contiguous node slots
        std::vector<int> internalKids, leafKids;
        for (int k : kids) (b.nodes[k].left < 0 ? leafKids : internalKids).push_back(k);
        n.childBase = (uint32_t)out.nodes.size();
        n.triBase = (uint32_t)out.leafOrder.size();
        const uint32_t childBase = n.childBase;
        out.nodes.resize(out.nodes.size() + internalKids.size());
        int slot = 0;
        uint32_t triOff = 0, intOff = 0;
        std::vector<std::pair<int, uint32_t>> recurse;
        auto quant = [&](const AABB& cb, int c) {
            const float org[3] = { n.px, n.py, n.pz };
            for (int a = 0; a < 3; a++)
            {
                float lo = std::floor((cb.lo[a] - org[a]) / scale[a]);
                float hi = std::ceil((cb.hi[a] - org[a]) / scale[a]);
                // guard against rounding of the division itself
                while (lo > 0 && org[a] + lo * scale[a] > cb.lo[a]) lo -= 1;
                while (hi < 255 && org[a] + hi * scale[a] < cb.hi[a]) hi += 1;
                lo = std::min(std::max(lo, 0.0f), 255.0f);
                hi = std::min(std::max(hi, 0.0f), 255.0f);
                n.qlo[a][c] = (uint8_t)lo;
                n.qhi[a][c] = (uint8_t)hi;
            }
        };
        for (int k : internalKids)
        {
            n.meta[slot] = (uint8_t)(0x20u | intOff);
            quant(b.nodes[k].box, slot);
            recurse.push_back({ k, childBase + intOff });
            intOff++; slot++;
        }
        for (int k : leafKids)
        {
            const BNode& c = b.nodes[k];
            n.meta[slot] = (uint8_t)((c.count << 6) | triOff);
            quant(c.box, slot);
            for (uint32_t i = 0; i < c.count; i++)
                out.leafOrder.push_back(b.order[c.first + i]);
            triOff += c.count; slot++;
        }
        out.nodes[outIdx] = n;
        const uint32_t pushed = (uint32_t)recurse.size();
        out.maxStack = std::max(out.maxStack, stackBelow + pushed);
        for (auto& r : recurse)
            emit_wide(b, r.first, r.second, out, depth + 1, stackBelow + pushed - 1);
    }

}

void build_bvh8(const float* wt, uint32_t total, BvhBuild& w)
{
    Builder b;
    b.triBox.resize(total); b.centroid.resize((size_t)total * 3); b.order.resize(total);
    for (uint32_t i = 0; i < total; i++)
    {
        const float* t = &wt[(size_t)i * 9];
        float p[3][3];
        for (int a = 0; a < 3; a++) { p[0][a] = t[a]; p[1][a] = t[a] + t[3 + a]; p[2][a] = t[a] + t[6 + a]; }
        AABB bx; bx.reset();
        for (int k = 0; k < 3; k++) bx.grow(p[k]);
        for (int a = 0; a < 3; a++)
        {
            const float pad = 4e-7f * std::max(std::max(std::fabs(bx.lo[a]), std::fabs(bx.hi[a])), 1.0f);
            bx.lo[a] -= pad; bx.hi[a] += pad;
            b.centroid[(size_t)i * 3 + a] = 0.5f * (bx.lo[a] + bx.hi[a]);
        }
        b.triBox[i] = bx;
        b.order[i] = i;
    }
    b.nodes.reserve((size_t)total * 2);
    b.build(0, total);
    w.nodes.clear(); w.leafOrder.clear(); w.maxDepth = 0; w.maxStack = 1;
    w.nodes.resize(1);
    w.leafOrder.reserve(total);
    emit_wide(b, 0, 0, w, 1, 0);
}
} // namespace zr
