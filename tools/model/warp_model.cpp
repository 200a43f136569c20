// warp_model.cpp — analysis tool (not product, not test): runs the kernel's per-ray code (trace_core.cuh, host build) over
// BASELINE config 2 and models how a warp schedule turns per-ray trip counts into warp instructions: lane efficiency of
// the 8x4 / 4x8 / 16x2 footprints, the share of warp trips by live lanes, what a CTA-level tail merge or a lane refill
// could save. Results are quoted in profiles/README.md.
//   python tools/model/dump_scene.py 12 && g++ -O2 -std=c++17 -ffp-contract=off -DYV_TEST_HOST_BUILD -o tools/model/_data/wm tools/model/warp_model.cpp && tools/model/_data/wm 8
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../yoxel-voxel_b200/csrc/trace_core.cuh"
using namespace yv;
struct St { U4 a[kMaxStack], b[kMaxStack]; void push(int sp,const U4&x,const U4&y){a[sp]=x;b[sp]=y;} void pop(int sp,U4&x,U4&y)const{x=a[sp];y=b[sp];} };
struct Fetch { const Rec*recs; 
  void node(uint32_t idx,bool,uint32_t&m,uint32_t&cb)const{m=recs[idx].masks;cb=recs[idx].child_base;}
  uint32_t child_index(uint32_t,uint32_t cb,uint32_t m,uint32_t c)const{return cb+(uint32_t)__builtin_popcount((m>>8)&((1u<<c)-1u));}
  uint32_t root_index()const{return 0;} };
int main(int argc,char**argv){
  FILE*f=fopen("tools/model/_data/recs.bin","rb"); fseek(f,0,SEEK_END); long n=ftell(f)/16; fseek(f,0,SEEK_SET);
  std::vector<Rec> recs(n); fread(recs.data(),16,n,f); fclose(f);
  float cam[9]; f=fopen("tools/model/_data/cam.bin","rb"); fread(cam,4,9,f); fclose(f);
  const int W=1920,H=1080; const float pos[3]={0.5f,0.5f,0.3f};
  int TW=argc>1?atoi(argv[1]):8, TH=32/TW;
  Fetch fetch{recs.data()};
  std::vector<int> trips(W*H,0);
  // per-trip kind per ray: store sequence? only counts: kind0 = trip ended in descend, 1 = pop, 2 = steps only (continue), 3=terminal
  long kinds[4]={0,0,0,0};
  for(int y=0;y<H;y++)for(int x=0;x<W;x++){
    float dx,dy,dz; primary_dir(cam,cam+3,cam+6,x,y,dx,dy,dz); dx=adjust_dir1(dx);dy=adjust_dir1(dy);dz=adjust_dir1(dz);
    LeanState s; St st; int t=0;
    if(lean_begin(s,fetch,true,pos[0],pos[1],pos[2],dx,dy,dz)){
      for(;;){ int sp0=s.sp; uint32_t idx0=s.idx; int r=lean_step<false>(s,fetch,st,false); t++;
        if(r!=kStepContinue){kinds[3]++;break;}
        if(s.idx!=idx0){ if(s.sp<sp0) kinds[1]++; else kinds[0]++; } else kinds[2]++; }
    }
    trips[y*W+x]=t;
  }
  // warp stats
  double sumL=0,sumMax=0; long warps=0; double hist[9]={0};
  // also: active lanes integrated over trips: for each warp, for trip k: lanes with L_i>k
  for(int ty=0;ty<H;ty+=TH)for(int tx=0;tx<W;tx+=TW){
    int mx=0; long s=0; for(int j=0;j<TH;j++)for(int i=0;i<TW;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W){int L=trips[yy*W+xx]; s+=L; mx=std::max(mx,L);}}
    // vote every 4 trips: warp runs ceil(mx/4)*4 trips
    int mx4=(mx+3)/4*4; sumL+=s; sumMax+=32.0*mx4; warps++;
    for(int k=0;k<mx4;k++){int act=0; for(int j=0;j<TH;j++)for(int i=0;i<TW;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W&&trips[yy*W+xx]>k)act++;} hist[(act+3)/4]+=1;}
  }
  long tot=0; for(int i=0;i<W*H;i++)tot+=trips[i];
  printf("tile %dx%d: trips/ray %.2f  trip-level lane efficiency (sumL/32*max) %.3f  warps %ld  warp-trips/warp %.1f\n",TW,TH,(double)tot/(W*H),sumL/sumMax,warps,sumMax/32/warps);
  printf("trip kinds: descend %.2f pop %.2f steps-only %.2f terminal %.2f per ray\n",kinds[0]/(double)(W*H),kinds[1]/(double)(W*H),kinds[2]/(double)(W*H),kinds[3]/(double)(W*H));
  double ht=0; for(int i=0;i<9;i++)ht+=hist[i]; printf("warp-trip share by active lanes (0,1-4,5-8,...,29-32): "); for(int i=0;i<9;i++)printf("%.3f ",hist[i]/ht); printf("\n");
  // CTA tail merge model: 16x8 tile = 4 warps 8x4; each warp runs (votes every 4 trips) until active<=T then parks; last warp adopts all parked
  if(TW==8) for(int T: {0,2,4,6,8}){
    double base=0, merged=0; 
    for(int ty=0;ty<H;ty+=8)for(int tx=0;tx<W;tx+=16){
      std::vector<int> rem; 
      for(int w=0;w<4;w++){ int wx=tx+(w&1)*8, wy=ty+(w>>1)*4; int L[32],n=0,mx=0; for(int j=0;j<4;j++)for(int i=0;i<8;i++){int yy=wy+j,xx=wx+i; int v=(yy<H&&xx<W)?trips[yy*W+xx]:0; L[n++]=v; mx=std::max(mx,v);} 
        int mx4=(mx+3)/4*4; base+=mx4;
        int k=0; for(;;k+=4){ int act=0; for(int i=0;i<32;i++) if(L[i]>k) act++; if(act<=T) break; }
        merged+=k; for(int i=0;i<32;i++) if(L[i]>k) rem.push_back(L[i]-k);
      }
      int mr=0; for(int v:rem) mr=std::max(mr,v); merged+=(mr+3)/4*4;
    }
    printf("T=%d: warp-trips base %.0f merged %.0f ratio %.3f\n",T,base,merged,merged/base);
  }
  // tail relaunch model: a warp that is down to <= T live lanes queues those pixels and exits; a second launch traces the
  // queued pixels from the start, 32 per warp in queue order
  if(TW==8) for(int T: {1,2,3,4,6,8}){
    double base=0, k1=0, k2=0; std::vector<int> q;
    for(int ty=0;ty<H;ty+=4)for(int tx=0;tx<W;tx+=8){
      int L[32],n=0,mx=0; for(int j=0;j<4;j++)for(int i=0;i<8;i++){int yy=ty+j,xx=tx+i; int v=(yy<H&&xx<W)?trips[yy*W+xx]:0; L[n++]=v; mx=std::max(mx,v);}
      base+=(mx+3)/4*4;
      int k=0; for(;;k+=4){ int act=0; for(int i=0;i<32;i++) if(L[i]>k) act++; if(act<=T) break; }
      k1+=k; for(int i=0;i<32;i++) if(L[i]>k) q.push_back(L[i]);
    }
    for(size_t i=0;i<q.size();i+=32){ int mx=0; for(size_t j=i;j<std::min(q.size(),i+32);j++) mx=std::max(mx,q[j]); k2+=(mx+3)/4*4; }
    printf("relaunch T=%d: queued %.1f%% of rays; warp-trips base %.0f -> %.0f + %.0f = ratio %.3f\n",T,100.0*q.size()/(W*H),base,k1,k2,(k1+k2)/base);
  }
  // persistent-refill model: one warp streams over 8x8 tiles (Morton order inside), refills idle lanes when active<=thr (checked every 4 trips)
  if(TW==8) for(int thr: {0,4,8,12,16,20,24,28}){
    // stream: tiles in raster order; sample every 7th row of tiles to keep it fast
    double warp_trips=0, lane_trips=0; 
    for(int ty=0;ty<H;ty+=8*4){ // one "warp" per sampled tile row
      std::vector<int> q; for(int tx=0;tx<W;tx+=8) for(int m=0;m<64;m++){int x=tx+((m&1)|((m>>1)&2)|((m>>2)&4)), y=ty+(((m>>1)&1)|((m>>2)&2)|((m>>3)&4)); if(x<W&&y<H) q.push_back(trips[y*W+x]);}
      size_t qi=0; int rem[32]={0}; 
      for(;;){ int act=0; for(int i=0;i<32;i++) if(rem[i]>0) act++;
        if(act<=thr || act==0){ for(int i=0;i<32&&qi<q.size();i++) if(rem[i]<=0){ rem[i]=q[qi++]; lane_trips+=rem[i]; } act=0; for(int i=0;i<32;i++) if(rem[i]>0) act++; if(act==0){ if(qi>=q.size()) break; else continue; } }
        for(int i=0;i<32;i++) rem[i]-=4; warp_trips+=4; }
    }
    printf("refill thr=%d: lane efficiency %.3f\n",thr,lane_trips/(32*warp_trips));
  }
  return 0; }
